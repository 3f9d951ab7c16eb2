"""Drop-in counterparts of the reference's `modules.py` classes that sit on the training hot path.

Same constructor signatures, forward signatures, initialisation distributions and state_dict keys as
the reference (so checkpoints written by utils.save_model, utils.py:645-673, load either way):

    SingleBVPNet   modules.py:430-481   (keys net.net.{i}.0.{weight,bias})
    FCBlock        modules.py:38-117
    BatchLinear    modules.py:12-26
    Sine           modules.py:29-35     (+ SineLayer = BatchLinear followed by Sine, the north-star's name
                                          for the anonymous MetaSequential pair built at modules.py:67-74)
    SimpleMLP      modules.py:256-312   (keys layers.{i}.{weight,bias})
    TonemapNet     modules.py:197-253   (keys layers_{r,g,b}.{0,2}.{weight,bias})
    LearnExps      modules.py:400-415   (key exps)
    LearnFocal     modules.py:378-397   (key focus)
    PosEncodingNeRF modules.py:537-580

Execution: `SingleBVPNet.forward` on the shipped atlas shape (2 -> 512 x4 -> 3|4, sine) runs the fused
tcgen05 chain kernels through the C ABI (forward, input/activation backward, weight gradients) and is
differentiable w.r.t. its parameters and its input coordinates like the reference (modules.py:460-464).
It raises on CPU tensors or when the native library is missing -- there is no fallback.  The small
per-pixel networks (SimpleMLP, TonemapNet, LearnExps, LearnFocal) are ordinary torch modules at module
level (used for warm-ups and rendering); inside the training step they are executed by the fused
per-pixel CUDA kernels of neucam_b200/step.py, which read their parameters in place.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import _gradmode, _lib, ops

HID = ops.HID


# ------------------------------------------------------------------------------------------------
# meta-module plumbing (torchmeta.modules: module.py:6-26, container.py:6-19, utils.py:4-11)
# ------------------------------------------------------------------------------------------------
def get_subdict(dictionary, key=None):
    if dictionary is None:
        return None
    if key is None or key == "":
        return dictionary
    prefix = key + "."
    return OrderedDict((k[len(prefix):], v) for k, v in dictionary.items() if k.startswith(prefix))


class MetaModule(nn.Module):
    """nn.Module whose forward accepts an optional `params` OrderedDict overriding its own parameters."""

    def meta_named_parameters(self, prefix="", recurse=True):
        return self.named_parameters(prefix=prefix, recurse=recurse)

    def meta_parameters(self, recurse=True):
        for _, p in self.meta_named_parameters(recurse=recurse):
            yield p


class MetaSequential(nn.Sequential, MetaModule):
    def forward(self, input, params=None):
        for name, module in self._modules.items():
            if isinstance(module, MetaModule):
                input = module(input, params=get_subdict(params, name))
            else:
                input = module(input)
        return input


class BatchLinear(nn.Linear, MetaModule):
    """Linear layer whose weights can be swapped per call (hypernetwork-style `params`)."""

    def forward(self, input, params=None):
        if params is None:
            params = OrderedDict(self.named_parameters())
        weight = params["weight"]
        bias = params.get("bias", None)
        out = input.matmul(weight.transpose(-1, -2))
        if bias is not None:
            out = out + bias.unsqueeze(-2)
        return out


class Sine(nn.Module):
    def forward(self, input):
        return torch.sin(30 * input)


class SineLayer(MetaSequential):
    """BatchLinear + Sine -- the pair the reference builds anonymously at modules.py:67-74."""

    def __init__(self, in_features, out_features):
        super().__init__(BatchLinear(in_features, out_features), Sine())


def sine_init(m):
    with torch.no_grad():
        if hasattr(m, "weight"):
            n = m.weight.size(-1)
            m.weight.uniform_(-np.sqrt(6 / n) / 30, np.sqrt(6 / n) / 30)


def first_layer_sine_init(m):
    with torch.no_grad():
        if hasattr(m, "weight"):
            n = m.weight.size(-1)
            m.weight.uniform_(-1 / n, 1 / n)


def init_weights_normal(m):
    if type(m) == BatchLinear or type(m) == nn.Linear:
        if hasattr(m, "weight"):
            nn.init.kaiming_normal_(m.weight, a=0.0, nonlinearity="relu", mode="fan_in")


# ------------------------------------------------------------------------------------------------
# fused atlas network
# ------------------------------------------------------------------------------------------------
class _SineMLPFunction(torch.autograd.Function):
    """y = MLP(uv) through the fused tcgen05 chain; backward = dgrad chain + split-K wgrad kernels."""

    @staticmethod
    def forward(ctx, opts, uv, w0, b0, w1, b1, w2, b2, w3, b3, w4, b4):
        grad_enabled, ctx.grad_targets = opts
        Q = uv.shape[0]
        dev = uv.device
        w16 = torch.empty((3 * HID, HID), dtype=torch.float16, device=dev)
        wT16 = torch.empty((3 * HID, HID), dtype=torch.float16, device=dev)
        w1c, w2c, w3c = w1.contiguous(), w2.contiguous(), w3.contiguous()
        ops.pack_hidden_weights(w1c, w2c, w3c, w16, wT16)
        b_hid = torch.stack((b1, b2, b3)).contiguous()
        # needs_input_grad ignores torch.no_grad(), and grad mode is always off inside forward(): the caller samples it
        need_grad = grad_enabled and any(ctx.needs_input_grad)
        uvc, w0c, b0c, w4c, b4c = uv.contiguous(), w0.contiguous(), b0.contiguous(), w4.contiguous(), b4.contiguous()
        if need_grad:
            act, cos, dz = ops.alloc_chain_workspace(Q, dev)
            y = ops.sine_mlp_fwd(uvc, w0c, b0c, w16, b_hid, w4c, b4c, act16=act, phase16=cos)
            w4T16 = torch.empty((HID, 64), dtype=torch.float16, device=dev)
            ops.pack_head_weights(w4c, w4T16)
            ctx.save_for_backward(uvc, w0c, b0c, w4c, wT16, act, cos, w4T16)
            ctx.dz = dz
        else:
            y = ops.sine_mlp_fwd(uvc, w0c, b0c, w16, b_hid, w4c, b4c)
        return y

    @staticmethod
    def backward(ctx, dy):
        uv, w0, b0, w4, wT16, act, cos, w4T16 = ctx.saved_tensors
        dz = ctx.dz
        dev = uv.device
        dy = dy.contiguous().float()
        tg = ctx.grad_targets
        # head bias gradient db4 += dy.sum(0) and the loss scale of max|dy| from one pass over dy
        db4 = tg[9] if tg is not None else torch.zeros(dy.shape[1], dtype=torch.float32, device=dev)
        _, scale = ops.relu_head_prep(dy, None, 0, db4)
        if tg is not None:
            # the step owns preallocated gradient buffers: every kernel below ACCUMULATES, so let them add straight
            # into param.grad and hand autograd nothing for the parameters
            dw0, db0, dws, dbs, dw4 = tg[0], tg[1], [tg[2], tg[4], tg[6]], [tg[3], tg[5], tg[7]], tg[8]
        else:
            dw0 = torch.zeros_like(w0)
            db0 = torch.zeros_like(b0)
            dws = [torch.zeros((HID, HID), dtype=torch.float32, device=dev) for _ in range(3)]
            dbs = [torch.zeros((HID,), dtype=torch.float32, device=dev) for _ in range(3)]
            dw4 = torch.zeros_like(w4)
        duv = ops.sine_mlp_bwd(uv, w0, b0, wT16, w4T16, dy, cos, scale, dz, dw0, db0)
        ops.sine_mlp_wgrad(dz, act, scale, dws, dbs)
        ops.sine_mlp_head_wgrad(cos, dy, dw4)
        if tg is not None:
            return (None, duv) + (None,) * 10
        return None, duv, dw0, db0, dws[0], dbs[0], dws[1], dbs[1], dws[2], dbs[2], dw4, db4


class FCBlock(MetaModule):
    """Fully connected network; with nonlinearity='sine' and the shipped atlas shape it runs fused."""

    def __init__(self, in_features, out_features, num_hidden_layers, hidden_features, outermost_linear=False,
                 nonlinearity="relu", weight_init=None):
        super().__init__()
        self.first_layer_init = None
        nls_and_inits = {"sine": (Sine(), sine_init, first_layer_sine_init),
                         "relu": (nn.ReLU(inplace=True), init_weights_normal, None)}
        if nonlinearity not in nls_and_inits:
            raise NotImplementedError(
                f"FCBlock nonlinearity '{nonlinearity}' is outside the NeuCam hot path (no shipped config uses it)")
        nl, nl_weight_init, first_layer_init = nls_and_inits[nonlinearity]
        self.nonlinearity = nonlinearity
        self.outermost_linear = outermost_linear
        self.dims = (in_features, hidden_features, num_hidden_layers, out_features)
        self.weight_init = weight_init if weight_init is not None else nl_weight_init
        net = [MetaSequential(BatchLinear(in_features, hidden_features), nl)]
        for _ in range(num_hidden_layers):
            net.append(MetaSequential(BatchLinear(hidden_features, hidden_features), nl))
        if outermost_linear:
            net.append(MetaSequential(BatchLinear(hidden_features, out_features)))
        else:
            net.append(MetaSequential(BatchLinear(hidden_features, out_features), nl))
        self.net = MetaSequential(*net)
        if self.weight_init is not None:
            self.net.apply(self.weight_init)
        if first_layer_init is not None:
            self.net[0].apply(first_layer_init)

    def fused_eligible(self):
        i, h, n, o = self.dims
        return (self.nonlinearity == "sine" and self.outermost_linear and i == 2 and h == HID and n == 3
                and o in (3, 4))

    def forward(self, coords, params=None, **kwargs):
        if params is None:
            params = OrderedDict(self.named_parameters())
        sub = get_subdict(params, "net")
        if not self.fused_eligible():
            raise NotImplementedError(
                "neucam_b200 executes the sine FCBlock shape every NeuCam config ships "
                "(in=2, hidden=512, 3 hidden layers, out 3|4, outermost_linear); got %s" % (self.dims,))
        if not coords.is_cuda:
            raise _lib.NeucamNativeError("FCBlock.forward needs CUDA tensors: the sm_100a kernels have no CPU fallback")
        lead = coords.shape[:-1]
        uv = coords.reshape(-1, 2).float()
        args = []
        for i in range(5):
            args += [sub[f"{i}.0.weight"], sub[f"{i}.0.bias"]]
        y = _SineMLPFunction.apply((torch.is_grad_enabled(), _gradmode.targets_of(args)), uv, *args)
        return y.reshape(*lead, y.shape[-1])


def _fcblock_forward_with_activations(self, coords, params=None, retain_grad=False):
    """FCBlock.forward_with_activations (modules.py:97-117): the model output AND every sublayer's output, as an
    OrderedDict keyed '<sublayer class>_<layer index>' after 'input'.  A DIAGNOSTIC entry point (its one caller in the
    reference is the activation-histogram summary, utils.py:46): it evaluates layer by layer with plain torch ops in
    fp32 on whatever device the input lives on, i.e. it shows what the reference would show -- not the fused kernels'
    fp16 operand tiles -- and is not part of the training path."""
    if params is None:
        params = OrderedDict(self.named_parameters())
    activations = OrderedDict()
    x = coords.clone().detach().requires_grad_(True)
    activations["input"] = x
    for i, layer in enumerate(self.net):
        subdict = get_subdict(params, "net.%d" % i)
        for j, sublayer in enumerate(layer):
            if isinstance(sublayer, BatchLinear):
                x = sublayer(x, params=get_subdict(subdict, "%d" % j))
            else:
                x = sublayer(x)
            if retain_grad:
                x.retain_grad()
            activations["_".join((str(sublayer.__class__), "%d" % i))] = x
    return activations


FCBlock.forward_with_activations = _fcblock_forward_with_activations


class SingleBVPNet(MetaModule):
    """The scene ("atlas") network: coords -> log radiance."""

    def __init__(self, out_features=1, type="sine", in_features=2, mode="mlp", hidden_features=256,
                 num_hidden_layers=3, **kwargs):
        super().__init__()
        if mode != "mlp":
            raise NotImplementedError("only mode='mlp' is on the NeuCam hot path (every shipped config uses it)")
        if kwargs.get("downsample", False):
            raise NotImplementedError("downsample=True is never used by the NeuCam configs")
        self.mode = mode
        self.net = FCBlock(in_features=in_features, out_features=out_features, num_hidden_layers=num_hidden_layers,
                           hidden_features=hidden_features, outermost_linear=True, nonlinearity=type)

    def forward(self, model_input, params=None):
        if params is None:
            params = OrderedDict(self.named_parameters())
        coords_org = model_input["coords"]
        if coords_org.shape[-1] == 3:
            coords_org = coords_org.clone().detach().requires_grad_(True)
        output = self.net(coords_org, get_subdict(params, "net"))
        return {"model_in": coords_org, "model_out": output}

    def forward_with_activations(self, model_input):
        """modules.py:477-481: {'model_in', 'model_out': (key, tensor) of the last sublayer, 'activations'}."""
        coords = model_input["coords"].clone().detach().requires_grad_(True)
        activations = self.net.forward_with_activations(coords)
        return {"model_in": coords, "model_out": activations.popitem(), "activations": activations}


# ------------------------------------------------------------------------------------------------
# small per-pixel networks
# ------------------------------------------------------------------------------------------------
class PosEncodingNeRF(nn.Module):
    """[c, sin(2^0 pi c_0), cos(2^0 pi c_0), sin(2^0 pi c_1), ...] (modules.py:567-580)."""

    def __init__(self, in_features, sidelength=None, fn_samples=None, use_nyquist=True, num_frequencies=None):
        super().__init__()
        self.in_features = in_features
        if in_features == 3:
            self.num_frequencies = 10
        elif in_features == 2:
            assert sidelength is not None
            if isinstance(sidelength, int):
                sidelength = (sidelength, sidelength)
            self.num_frequencies = 4
            if use_nyquist:
                self.num_frequencies = self.get_num_frequencies_nyquist(min(sidelength[0], sidelength[1]))
        elif in_features == 1:
            assert fn_samples is not None
            self.num_frequencies = 4
            if use_nyquist:
                self.num_frequencies = self.get_num_frequencies_nyquist(fn_samples)
        if num_frequencies is not None:
            self.num_frequencies = num_frequencies
        self.out_dim = in_features + 2 * in_features * self.num_frequencies

    @staticmethod
    def get_num_frequencies_nyquist(samples):
        nyquist_rate = 1 / (2 * (2 * 1 / samples))
        return int(math.floor(math.log(nyquist_rate, 2)))

    def forward(self, coords):
        coords = coords.view(coords.shape[0], -1, self.in_features)
        feats = [coords]
        for i in range(self.num_frequencies):
            arg = (2 ** i) * np.pi * coords                       # [..., in]
            sc = torch.stack((torch.sin(arg), torch.cos(arg)), -1)  # [..., in, 2] -> sin_j, cos_j interleaved
            feats.append(sc.reshape(*coords.shape[:-1], 2 * self.in_features))
        return torch.cat(feats, -1).reshape(coords.shape[0], -1, self.out_dim)


class SimpleMLP(nn.Module):
    """ReLU MLP used as blur-kernel / uv-mapping / alpha network."""

    def __init__(self, hidden_dim=256, num_layer=8, in_dim=3, out_dim=2, use_encoding=False, num_frequencies=None,
                 skip_layer=[4], nonlinear="tanh") -> None:
        super().__init__()
        self.use_encoding = use_encoding
        self.skip_layer = skip_layer
        self.last_nonlinear = nonlinear
        self.num_layer = num_layer
        self.layers = nn.ModuleList()
        self.in_dim = in_dim
        if self.use_encoding:
            self.positional_encoding = PosEncodingNeRF(in_features=in_dim, num_frequencies=num_frequencies)
            in_dim = self.positional_encoding.out_dim
        input_ch = in_dim
        for i in range(num_layer):
            if i in skip_layer:
                input_ch += in_dim
            output_ch = out_dim if i == num_layer - 1 else hidden_dim
            self.layers.append(nn.Linear(input_ch, output_ch))
            input_ch = output_ch

    def forward(self, x):
        if x.is_cuda and self.layers[0].out_features == 256:
            # the uv-mapping / alpha networks: tcgen05 chain kernels (forward, input gradient, weight gradients)
            from . import relu_mlp
            if relu_mlp.eligible(self):
                return relu_mlp.fused_forward(self, x)
        if self.use_encoding:
            x = self.positional_encoding(x).view(-1, self.positional_encoding.out_dim)
        else:
            x = x.reshape(-1, self.in_dim)
        skip_in = x.detach()
        for i, layer in enumerate(self.layers):
            if i in self.skip_layer:
                x = torch.cat((x, skip_in), -1)
            x = layer(x)
            if i != self.num_layer - 1:
                x = F.relu(x)
        kind = self.last_nonlinear
        if kind == "sigmoid":
            return torch.sigmoid(x)
        if kind == "softmax":
            return torch.softmax(x, -1)
        if kind == "tanh":
            return torch.tanh(x)
        if kind == "both":
            k = x.shape[-1] // 3
            return torch.cat([torch.softmax(x[..., :k], -1), torch.tanh(x[..., k:])], -1)
        if kind == "none":
            return x
        raise Exception("Unknown activation functions.", kind)


class _TonemapFunction(torch.autograd.Function):
    """TonemapNet.forward through the standalone CRF kernels (csrc/tonemap.cu): warm-up and rendering callers."""

    @staticmethod
    def forward(ctx, x, expo, *params):
        ps = [p.contiguous() for p in params]
        xc, ec = x.contiguous().float(), expo.reshape(-1).contiguous().float()
        ctx.save_for_backward(xc, ec, *ps)
        ctx.expo_shape = expo.shape
        return ops.tm_fwd(ps, xc, ec)

    @staticmethod
    def backward(ctx, dldr):
        xc, ec, *ps = ctx.saved_tensors
        grads = [torch.zeros_like(p) for p in ps]
        dx, dexpo = ops.tm_bwd(ps, xc, ec, dldr.contiguous().float(), grads, want_dexpo=ctx.needs_input_grad[1])
        return (dx, dexpo.view(ctx.expo_shape) if dexpo is not None else None, *grads)


class TonemapNet(nn.Module):
    """Per-channel 1 -> W -> 1 ReLU MLP camera response function on log exposure."""

    def __init__(self, W=128):
        super().__init__()
        self.layers_r = nn.Sequential(nn.Linear(1, W), nn.ReLU(), nn.Linear(W, 1))
        self.layers_g = nn.Sequential(nn.Linear(1, W), nn.ReLU(), nn.Linear(W, 1))
        self.layers_b = nn.Sequential(nn.Linear(1, W), nn.ReLU(), nn.Linear(W, 1))

    def forward(self, x, input_exps, noise=None, output_grads=False):
        if noise is not None:
            raise NotImplementedError("the noise branch (opt.denoise) is not used by any shipped config")
        if x.is_cuda and not output_grads and self.layers_r[0].out_features == ops.TM_WIDTH and x.dim() == 2:
            # CUDA: the CRF kernels (csrc/tonemap.cu); exposure [M,1] or [M]
            return _TonemapFunction.apply(x, input_exps.expand(x.shape[0], 1) if input_exps.dim() == 2 else input_exps,
                                          *self.parameters())
        lnx = x + math.log(2.0) * input_exps
        ldr = torch.cat([self.layers_r(lnx[:, 0:1]), self.layers_g(lnx[:, 1:2]), self.layers_b(lnx[:, 2:3])], -1)
        ldr = torch.tanh(ldr)
        if output_grads:
            grad = torch.autograd.grad(ldr.mean(), lnx, retain_graph=True)[0].min()
            return ldr, grad
        return ldr


class LearnFocal(nn.Module):
    """Per-frame focus scalar."""

    def __init__(self, focal_list=[-1., -1., 0., 1., 1.], requires_grad=False):
        super().__init__()
        if isinstance(focal_list, list):
            focus = torch.tensor(focal_list, requires_grad=requires_grad)
        else:
            focus = torch.zeros(focal_list, requires_grad=True)
        self.focus = nn.Parameter(focus)

    def forward(self):
        return self.focus

    def get_minmax(self):
        return torch.stack((self.focus.min(), self.focus.max()))

    def get_mid(self):
        return torch.sort(self.focus)[0][len(self.focus) // 2]


class LearnExps(nn.Module):
    """Per-frame exposure, 3 tanh(e) in [-3, 3] stops."""

    def __init__(self, num=18):
        super().__init__()
        self.exps = nn.Parameter(torch.zeros(num, requires_grad=True))

    def forward(self):
        return torch.tanh(self.exps) * 3

    def get_minmax(self):
        return torch.stack((self.exps.min(), self.exps.max()))

    def get_mid(self):
        return torch.sort(self.exps)[0][len(self.exps) // 2]
