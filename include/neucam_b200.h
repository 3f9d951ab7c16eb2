/* neucam_b200.h -- C ABI of libneucam_b200.so, the sm_100a implementation of NeuCam's training hot
 * path (reference: xhuangcv/neucam, training.py:95-271 over modules.py).
 *
 * The reference has no FFI of its own: its boundary is the Python nn.Module surface of modules.py.
 * Every entry point below therefore names the reference statements it replaces; the Python side
 * (neucam_b200/modules.py, neucam_b200/step.py) binds them with ctypes -- see INTEGRATION.md.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host; the library never allocates,
 *     frees or retains them (PyTorch owns every buffer);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - return value: 0 = ok, >0 = cudaError_t, <0 = NEUCAM_ERR_*;
 *   - there is no CPU fallback: a non-sm_100 device or a missing driver entry point is an error;
 *   - NO FLOAT ATOMICS: every entry point that reduces over thread blocks (gradients of shared parameters, loss
 *     sums) takes a scratch workspace `ws` of at least neucam_reduce_ws_bytes(kind, ...) bytes, writes one partial
 *     record per block there and adds the records up in block order in a second kernel.  Same binary + same
 *     inputs = bit-identical outputs.  The workspace is scratch: its contents are dead when the stream reaches the
 *     end of the call, so calls on ONE stream can share it; concurrent streams need their own.
 *   - global state: a cache of encoded TMA descriptors keyed by (device, pointer, geometry), the per-device
 *     "dynamic shared memory attribute set" marks, and a launch counter (neucam_launch_count).  Nothing else.
 */
#ifndef NEUCAM_B200_H_
#define NEUCAM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NEUCAM_ABI_VERSION 8

#define NEUCAM_ERR_BAD_ARG (-1)
#define NEUCAM_ERR_NO_DRIVER (-2) /* cuTensorMapEncodeTiled entry point unavailable */
#define NEUCAM_ERR_TMAP (-3)      /* tensor-map encoding rejected */
#define NEUCAM_ERR_ARCH (-4)      /* device is not compute capability 10.x */
#define NEUCAM_ERR_WORKSPACE (-5) /* reduction workspace missing or smaller than neucam_reduce_ws_bytes says */

/* ABI version of this header; bumped on any signature change. */
int neucam_abi_version(void);

/* 0 if the current device can run the library (compute capability 10.x), NEUCAM_ERR_ARCH otherwise. */
int neucam_check_device(void);

/* Bytes of reduction workspace an entry point needs on the current device.  kind = NEUCAM_WS_*;
 * n / aux: SINE_BWD, WGRAD (both wgrad entry points), HEAD_WGRAD (both): unused; PSF_CRF and BLUR_BWD: n = B pixels,
 * aux = N frames; RIGIDITY: n = B; BLEND: n = K * B; TM: n = M samples. */
#define NEUCAM_WS_SINE_BWD 0
#define NEUCAM_WS_WGRAD 1
#define NEUCAM_WS_HEAD_WGRAD 2
#define NEUCAM_WS_PSF_CRF 3
#define NEUCAM_WS_BLUR_BWD 4
#define NEUCAM_WS_RIGIDITY 5
#define NEUCAM_WS_BLEND 6
int neucam_reduce_ws_bytes(int kind, int64_t n, int aux, int64_t* bytes_out);

/* Number of CUDA kernels this library has launched so far in this process (bench.py's `gpu_launches`). */
int neucam_launch_count(int64_t* count_out);

/* ---- scene sine-MLP (replaces SingleBVPNet.forward -> FCBlock.forward -> BatchLinear/Sine chain,
 * modules.py:455-475, 90-95, 17-35, for the shipped shape 2 -> 512 x4 -> out_dim, and its autograd
 * backward w.r.t. activations and input coordinates, training.py:270) -------------------------
 *
 * neucam_sine_mlp_fwd:  y[Q,out_dim] = MLP(uv[Q,2]).
 *   w0 [512,2], b0 [512], b_hid [3,512], w4 [out_dim,512], b4 [out_dim] : fp32 master parameters.
 *   w_hid16 : fp16 [3*512,512], the three hidden weight matrices stacked, each row-major [out,in]
 *             (kept in sync with the fp32 masters by neucam_pack_hidden_weights).
 *   act16   : fp16 [3,Q,512] out, a1..a3 (inputs of layers 1..3, the weight-gradient operands), or NULL for inference.
 *             a4 is not stored: neucam_sine_mlp_head_wgrad regenerates it from the phase of z3.
 *   phase16 : 16-bit [3, ceil(Q/128)*128*512] out, the phase of 30 z_l (l=1..3) in revolutions * 65536,
 *             in the library's tile-blocked layout (opaque; only neucam_sine_mlp_bwd reads it, to
 *             regenerate cos(30 z_l)); NULL iff act16 is NULL.
 */
int neucam_sine_mlp_fwd(const float* uv, int64_t Q, const float* w0, const float* b0, const void* w_hid16,
                        const float* b_hid, const float* w4, const float* b4, int out_dim, float* y,
                        void* act16, void* phase16, void* stream);

/* neucam_sine_mlp_bwd:  given dy[Q,out_dim] computes duv[Q,2], ACCUMULATES the first layer's
 *   gradients into dw0 [512,2] / db0 [512] and writes dz16 = fp16 [3,Q,512] holding
 *   (*scale_dev) * dL/dz_l for l = 1..3 (consumed by neucam_sine_mlp_wgrad).
 *   wT_hid16 : fp16 [3*512,512], 30 * the hidden weights stacked and TRANSPOSED ([in,out] row-major).
 *   w4T16    : fp16 [512,64], the head weights transposed and zero-padded (neucam_pack_head_weights): the head
 *              dgrad dy W4 runs as one K = 16 tensor-core step.
 *   scale_dev: device pointer to one float, a power-of-two loss scale chosen so that the scaled
 *              gradients sit in fp16's normal range (neucam_b200/step.py derives it from max|dy|).
 */
int neucam_sine_mlp_bwd(const float* uv, int64_t Q, const float* w0, const float* b0, const void* wT_hid16,
                        const void* w4T16, int out_dim, const float* dy, const void* phase16,
                        const float* scale_dev, void* dz16, float* duv, float* dw0, float* db0, void* ws,
                        int64_t ws_bytes, void* stream);

/* neucam_sine_mlp_wgrad: ACCUMULATES dW_l [512,512] += dz_l^T a_l / scale and db_l [512] += colsum(dz_l) / scale
 *   for the hidden layers l = 1..3 (autograd MmBackward/sum of BatchLinear.forward, modules.py:24-25).
 *   dz16 = [3,Q,512] from neucam_sine_mlp_bwd, act16 = [3,Q,512] from neucam_sine_mlp_fwd. */
int neucam_sine_mlp_wgrad(const void* dz16, const void* act16, int64_t Q, const float* scale_dev, float* dw1,
                          float* db1, float* dw2, float* db2, float* dw3, float* db3, void* ws, int64_t ws_bytes,
                          void* stream);

/* neucam_sine_mlp_head_wgrad: ACCUMULATES dW4 [out_dim,512] += dy^T a4 with a4 = sin(30 z3) regenerated from
 *   phase3_16 = the third layer's slice of phase16 (phase16 + 2 * ceil(Q/128)*128*512 elements; whole 128-row tiles are
 *   read, as the forward wrote them).  out_dim 3 or 4; phase3_16 and dy 16-byte aligned (cp.async.bulk sources).
 *   Workspace: NEUCAM_WS_HEAD_WGRAD. */
int neucam_sine_mlp_head_wgrad(const void* phase3_16, const float* dy, int out_dim, int64_t Q, float* dw4, void* ws,
                               int64_t ws_bytes, void* stream);

/* ---- 256-wide ReLU MLPs: the uv-mapping and alpha networks of the layered configs --------------------------
 * Replaces SimpleMLP.forward (modules.py:282-306) as called at training.py:131-137 (uv_b / uv_f), 148-150
 * (alpha) and utils.py:418-419 (rigidity loss), and the autograd backward of those calls (training.py:270).
 *
 * The network input is handed over ENCODED: feat16 [Q,64] fp16 = raw coordinates + NeRF positional encoding
 * (modules.py:567-580) in any column order the caller likes, zero padded to 64; weights that multiply the
 * encoded input are given in the same column order.  A network with H+1 Linear layers is H tensor-core GEMMs
 * (first layer, hidden layers, with the reference's skip connection modules.py:299-300 as an extra K-block on
 * the encoded input) followed by the out_dim <= 4 output layer and its nonlinearity on CUDA cores.
 *
 * The forward runs SPLIT-fp16 (hi + lo pairs for activations and weights, three MMAs per product) because a ReLU
 * network's gradient is discontinuous in its pre-activations; the backward splits the transposed weights (their
 * rounding is common to all samples).  The "lo" parts of the per-sample tensors (act_lo16, dz_lo16) are OPTIONAL
 * everywhere: NULL = single fp16 (their rounding averages out over the samples); see csrc/relu_chain.cu.
 *
 * neucam_relu_mlp_fwd: w_stages [S*256,64] fp16 = for every GEMM in order its 256(out) x 64(in) K-blocks (hi
 *   and lo parts of the weights as separate stages); stage_begin[n_gemm+1] delimits them; a_src[S] says what each
 *   K-block multiplies (0..3 = that 64-column block of the previous activation's hi part, 4 = the encoded input)
 *   and a_src2[S] whether it also multiplies the lo part of a block (5..8, else 255).  The activation of the
 *   kernel's operand tile is hi = fp16(h), lo = fp16(h - hi).  bias [n_gemm,256]; w_last [out_dim,256], b_last fp32;
 *   out_kind 0 none / 1 tanh / 2 sigmoid.  Writes y [Q,out_dim] fp32 and, when act16 != NULL, the activations
 *   h_0..h_{H-1} as [H,Q,256] fp16 hi (act16) and lo (act_lo16) parts plus gates [H,Q,8] uint32 (bit u of a row:
 *   hidden unit u is open), which is all the backward chain reads of the forward.
 * neucam_relu_mlp_bwd: dpre [Q,out_dim] fp32 = dL/d(output-layer pre-activation); wT_stages = for layer H-1
 *   down to 1 (then layer 0 if dfeat != NULL) the four 256(in, zero padded) x 64(out) K-blocks of the TRANSPOSED
 *   weight, each as a hi stage then a lo stage (skip columns excluded: the reference detaches the skip input, modules.py:287).  Writes dz16 / dz_lo16
 *   [H,Q,256] fp16 (hi / lo parts, times *scale_dev, a power of two) and dfeat [Q,64] fp32 (unscaled) when asked.
 * neucam_relu_mlp_wgrad / _head_wgrad: ACCUMULATE the weight and bias gradients (divided by *scale_dev) from
 *   dz16 / act16 / feat16: dw_first [256,64] and dw_skip[l] [256,64] in encoded-input column order, dw[l] with row
 *   pitch ldw[l] for l >= 1 (its first 256 columns), db[l] [256]; dw_last [out_dim,256] += dpre^T h_{H-1}. */
int neucam_relu_mlp_fwd(const void* feat16, int64_t Q, const void* w_stages, int n_gemm, const int* stage_begin,
                        const uint8_t* a_src, const uint8_t* a_src2, const float* bias, const float* w_last,
                        const float* b_last, int out_dim, int out_kind, float* y, void* act16, void* act_lo16,
                        void* gates, void* stream);
int neucam_relu_mlp_bwd(const float* dpre, int64_t Q, const void* wT_stages, int H, const float* w_last, int out_dim,
                        const void* gates, const float* scale_dev, void* dz16, void* dz_lo16, float* dfeat,
                        void* stream);
int neucam_relu_mlp_wgrad(const void* dz16, const void* dz_lo16, const void* act16, const void* act_lo16,
                          const void* feat16, int64_t Q, int H, const float* scale_dev, float* dw_first,
                          float* const* dw, const int* ldw, float* const* db, float* const* dw_skip, void* ws,
                          int64_t ws_bytes, void* stream);
int neucam_relu_mlp_head_wgrad(const void* h16, const void* h_lo16, const float* dpre, int out_dim, int64_t Q,
                               float* dw_last, void* ws, int64_t ws_bytes, void* stream);

/* Helpers around the chain (csrc/relu_aux.cu).
 * neucam_relu_mlp_encode: x [Q,3] fp32 -> feat16 [Q,64] fp16 = [ hi(E) | lo(E) | 0 ], E = 3 + 6*num_freq, the
 *   features in PosEncodingNeRF's order (modules.py:567-580; num_freq = 0: use_encoding=False).
 * neucam_relu_mlp_encode_bwd: dfeat [Q,64] fp32 (from neucam_relu_mlp_bwd) -> dx [Q,3] through the encoding.
 * neucam_relu_mlp_pack: fp32 Linear weights w[l] ([out,in_l], row pitch ld[l]) -> w_stages for the forward plan
 *   (per stage: st_layer, st_kb 0..3 | 4 = encoded-input block, st_lo 0 hi / 1 lo) and wT_stages in
 *   neucam_relu_mlp_bwd's fixed order (NULL to skip). */
int neucam_relu_mlp_encode(const float* x, int64_t Q, int num_freq, void* feat16, void* stream);
/* neucam_relu_mlp_head_prep: the pointwise front of a ReLU MLP's backward in one pass over [Q,out_dim]: dpre = dy *
 *   d(out)/d(pre) from the forward output y (out_kind 1 tanh: 1 - y^2, 2 sigmoid: y (1 - y); 0: dpre = dy, in which
 *   case y and dpre are not touched and may be NULL -- the scene network's linear head uses it that way); ACCUMULATES
 *   db_last [out_dim] += sum_r dpre[r,:]; writes *scale_out = the power-of-two loss scale of max|dpre| (as
 *   neucam_loss_scale with `target`).  scratch_bits: 4 bytes.  Workspace: NEUCAM_WS_HEAD_PREP. */
int neucam_relu_mlp_head_prep(const float* dy, const float* y, int out_dim, int out_kind, int64_t Q, float target,
                              float* dpre, float* db_last, uint32_t* scratch_bits, float* scale_out, void* ws,
                              int64_t ws_bytes, void* stream);
int neucam_relu_mlp_encode_bwd(const float* x, const float* dfeat, int64_t Q, int num_freq, float* dx, void* stream);
int neucam_relu_mlp_pack(const float* const* w, const int* ld, int L, int E, int S, const uint8_t* st_layer,
                         const uint8_t* st_kb, const uint8_t* st_lo, int with_input_grad, void* w_stages,
                         void* wT_stages, void* stream);

/* ---- per-pixel stages around the scene MLP ------------------------------------------------------------
 * Layout conventions: a step samples B pixels; every pixel has K taps (K = patch_size = 9 with the blur
 * model, 1 without); per-tap tensors are tap-major [K,B,...] exactly like the reference's coords patch
 * (utils.py:587-612), so sample q = k*B + p.  Per-pixel feature scratch is feature-major ([F,B]).
 *
 * neucam_blur_fwd (training.py:101-121: utils.get_input_patch + LearnFocal gather + SimpleMLP blur net
 *   (modules.py:282-306, nonlinear='both') + defocus offsets + centre-tap reset):
 *   coords [B,3] (t,y,x), coord_idx [B] int64, focus [N]; blur_params = HOST array of 8 device pointers
 *   layers.0.weight [64,3], layers.0.bias, layers.1.weight [64,64], .bias, layers.2.weight, .bias,
 *   layers.3.weight [27,64], .bias.   Outputs: uv [K,B,2] (y,x of every tap), kout [27,B]
 *   (softmax weights | tanh y offsets | tanh x offsets), hid [3*64,B] post-ReLU activations (for bwd). */
int neucam_blur_fwd(const float* coords, const int64_t* coord_idx, const float* focus,
                    const float* const* blur_params, int B, int K, int N, int H, int W, float offset_size,
                    float* uv, float* kout, float* hid, void* stream);

/* neucam_blur_bwd: autograd backward of neucam_blur_fwd. dkw [K,B] = dL/d(kernel weight), duv [K,B,2] =
 *   dL/d(tap coords).  ACCUMULATES into blur_grads (HOST array of 8 device pointers, same order as
 *   blur_params) and dfocus [N]. */
int neucam_blur_bwd(const float* coords, const int64_t* coord_idx, const float* focus,
                    const float* const* blur_params, float* const* blur_grads, float* dfocus, int B, int K, int N,
                    int H, int W, float offset_size, const float* kout, const float* hid, const float* dkw,
                    const float* duv, void* ws, int64_t ws_bytes, void* stream);

/* neucam_psf_crf_loss (training.py:157-188, 244-245 forward AND backward in one pass): PSF-weighted tap sum,
 *   exposure dt = 3 tanh(exps[frame]) (LearnExps, modules.py:407-408) or fixed_exps [B], TonemapNet
 *   (modules.py:219-247; tm_params = HOST array of 12 device pointers: for r,g,b: 0.weight [128,1], 0.bias [128],
 *   2.weight [1,128], 2.bias [1]), image_mse (loss_functions.py:8-12) optionally with the random-gamma
 *   augmentation (utils.py:615-619; gamma_dev = device scalar holding the step's gamma or NULL, gamma_w [B]) and mask weights (mask_w [B] or NULL), and the
 *   white-balance term |tm(0,0) - fixed_value| weighted by ue_weight.
 *   global_count = 3 * (pixels of the GLOBAL batch) normalises the mean.
 *   Outputs: dy [K,B,3], dkw [K,B]; ACCUMULATES tm_grads (12 pointers), dexps [N], db4 [3] (atlas output bias), sums[0] += sum of squared
 *   errors, sums[1] += ue_loss; atomically maxes |dy| into dy_absmax_bits and then writes
 *   scale_out = 2^floor(log2(scale_target / max|dy|)) (the loss scale of neucam_sine_mlp_bwd).
 *   ldr_out [B,3] optional (tone-mapped prediction).  tm_width = hidden width of the TonemapNet the 12 pointers
 *   belong to: must be 128 (TonemapNet's default and the only width any config uses), anything else is rejected. */
int neucam_psf_crf_loss(const float* y, const float* kout, const int64_t* coord_idx, const float* exps,
                        const float* fixed_exps, const float* gt, const float* gamma_w, const float* mask_w,
                        const float* const* tm_params, float* const* tm_grads, int B, int K, int N, int H, int W,
                        int has_blur, const float* gamma_dev, float fixed_value, float ue_weight,
                        int64_t global_count, float* dy, float* dkw, float* dexps, float* db4, float* sums,
                        uint32_t* dy_absmax_bits, float scale_target, float* scale_out, float* ldr_out,
                        int tm_width, void* ws, int64_t ws_bytes, void* stream);

/* neucam_rigidity_loss (utils.get_rigidity_loss, utils.py:402-454, as called at training.py:203-217; value and
 *   gradient in one pass): uv_ref [B,2] = the uv mapping at the step's pixels, uv_moved [2,B,2] = at the pixel one row
 *   up / one column left.  ku = (H-1)/2 / uv_mapping_scale, kv = (W-1)/2 / uv_mapping_scale (the reference scales both
 *   u differences by H-1 and both v differences by W-1).  *coef_dev (device scalar) = loss weight / B (x data-parallel
 *   share x mean mask weight).  loss_sum[0] += coef * sum L_i; g_ref [B,2], g_moved [2,B,2] = coef * gradients.
 * neucam_loss_scale: scale_out[0] = 2^floor(log2(target / max|x|)) clamped to [2^-20, 2^64] (1 if max|x| is 0 or not
 *   finite): the loss scale of the fp16 backward chains, computed on the device. */
int neucam_rigidity_loss(const float* uv_ref, const float* uv_moved, int B, float ku, float kv, const float* coef_dev,
                         float* loss_sum, float* g_ref, float* g_moved, void* ws, int64_t ws_bytes, void* stream);
int neucam_loss_scale(const float* x, int64_t n, float target, uint32_t* scratch_bits, float* scale_out, void* stream);

/* neucam_layer_blend_fwd / _bwd (training.py:143-153 layer blend with the alpha MLP's output; 220-238 sparsity and
 *   alpha losses): back, front, out [K,B,3] (log radiance per tap), alpha [K,B]; a_ref = alpha of the centre tap.
 *   sums[0] += c1 * sum_{k,b} (1-a_ref)^2 sum_c exp(front)^2, sums[1] += c2 * sum_b -1/(log(a_ref+1e-24)+log(1-a_ref+1e-24)),
 *   sums[2] += c3 * sum_b a_ref  (the caller folds weights, means and the data-parallel share into c1..c3; 0 = off).
 *   e2 [K,B] is scratch handed from forward to backward.  The backward takes dout = dL/d out and the device scalar
 *   *g_reg = dL/d(the three sums) and writes d_back, d_front [K,B,3], d_alpha [K,B]. */
int neucam_layer_blend_fwd(const float* back, const float* front, const float* alpha, int K, int B, float c1, float c2,
                           float c3, float* out, float* e2, float* sums, void* ws, int64_t ws_bytes, void* stream);
int neucam_layer_blend_bwd(const float* dout, const float* back, const float* front, const float* alpha,
                           const float* e2, const float* g_reg, int K, int B, float c1, float c2, float c3,
                           float* d_back, float* d_front, float* d_alpha, void* stream);

/* neucam_tm_fwd / neucam_tm_bwd: TonemapNet.forward (modules.py:219-247, noise=None) on its own, for the callers around
 *   the step that evaluate the CRF alone -- utils.warm_tm (utils.py:265-282) and full-frame rendering
 *   (utils.write_video_summary, render_video.py).  x [M,3] log radiance, expo [M] exposure in stops; ldr [M,3] =
 *   tanh(CRF(x + ln2 * expo)).  The backward takes dldr [M,3], writes dx [M,3] and (optionally) dexpo [M], and
 *   ACCUMULATES the 12 parameter gradients (workspace: M/128 records of 1156 floats = NEUCAM_WS_TM with n = M). */
#define NEUCAM_WS_TM 7
#define NEUCAM_WS_HEAD_PREP 8 /* neucam_relu_mlp_head_prep (n, aux unused) */
int neucam_tm_fwd(const float* const* tm_params, int tm_width, const float* x, const float* expo, int64_t M,
                  float* ldr, void* stream);
int neucam_tm_bwd(const float* const* tm_params, int tm_width, const float* x, const float* expo, const float* dldr,
                  int64_t M, float* dx, float* dexpo, float* const* tm_grads, void* ws, int64_t ws_bytes,
                  void* stream);

/* neucam_tm_min_slope: the VALUE-ONLY CRF gradient regulariser of training.py:248-252 (TonemapNet.forward with
 *   output_grads=True, modules.py:249-251): min over M points of d mean(ldr)/d lnx for rad [M,3], expo [M,1].
 *   *min_bits receives an order-preserving uint32 encoding of the float minimum (decode: b & 0x80000000 ?
 *   b & 0x7fffffff : ~b).  grad_loss = relu(-min) * 1e6 contributes no gradient in the reference. */
int neucam_tm_min_slope(const float* const* tm_params, const float* rad, const float* expo, int M,
                        uint32_t* min_bits, void* stream);

/* neucam_sample_batch (replaces dataio.Implicit3DWrapper.__getitem__, dataio.py:356-393,
 *   for a video resident in HBM): data [N*H*W,3] (= img*2-1), optional gamma_w_all / mask_w_all [N*H*W].
 *   Draws B pixel indices with replacement (Philox-4x32-10, key = seed, counter = (pixel, step)) or uses
 *   idx_in [B] when given (index injection for parity runs); writes coords [B,3] (get_mgrid formula,
 *   dataio.py:24-45), coord_idx [B], img [B,3] and the gathered per-pixel weights. */
int neucam_sample_batch(const float* data, const float* gamma_w_all, const float* mask_w_all, const int64_t* idx_in,
                        int N, int H, int W, int B, uint64_t seed, uint64_t step, float* coords, int64_t* coord_idx,
                        float* img, float* gamma_w, float* mask_w, void* stream);

/* ---- optimizer (torch.optim.Adam as constructed at training.py:51) --------------------------------------
 * state = device float[4]: {step count, 1-beta1^t, 1-beta2^t, unused}.  neucam_adam_tick advances it once
 * per step; neucam_adam_step then updates one flat range (call once per learning-rate group).
 * grad_scale multiplies the gradient on read (1/world_size after a sum all-reduce, else 1). */
int neucam_adam_tick(float* state, float beta1, float beta2, void* stream);
int neucam_adam_step(float* p, const float* g, float* m, float* v, int64_t n, const float* state, float lr,
                     float beta1, float beta2, float eps, float grad_scale, void* stream);

/* Re-emits the fp16 operand copies of the three atlas hidden matrices (fp32 [512,512] each):
 * w16 = [3*512,512] row-major [out,in]; wT16 = [3*512,512] = 30 * W transposed ([in,out]; the dgrad chain's
 * 30*cos(.) factor is folded into the operand). */
int neucam_pack_hidden_weights(const float* w1, const float* w2, const float* w3, void* w16, void* wT16,
                               void* stream);
/* w4 fp32 [out_dim,512] -> w4T16 fp16 [512,64] (row j = W4[0..out_dim-1][j], zero padded). */
int neucam_pack_head_weights(const float* w4, int out_dim, void* w4T16, void* stream);

/* ---- the whole static step from ONE call (SURVEY.md section 8b "minimum set") ---------------------------------
 * neucam_step_fwd_bwd runs training.py:95-270 for the static multi-focus / multi-exposure model set (atlas_b +
 * blur + focal + exp + tm; K = 9 taps with the blur model, K = 1 without): blur taps -> scene MLP forward -> PSF sum +
 * exposure + CRF + image_mse (+ gamma / mask) + ue_loss and their backward -> scene MLP backward -> weight gradients
 * of every network, i.e. exactly the sequence neucam_blur_fwd, neucam_sine_mlp_fwd, neucam_psf_crf_loss,
 * neucam_sine_mlp_bwd, neucam_blur_bwd, neucam_sine_mlp_wgrad, neucam_sine_mlp_head_wgrad.  Every pointer is a
 * device pointer owned by the caller; gradients are ACCUMULATED into the g_* tensors after `zero_ptr[0 .. zero_bytes)`
 * (normally the flat gradient buffer all g_* live in) and `sums` have been cleared.  Outputs: gradients, sums[0] =
 * sum of squared errors, sums[1] = ue_loss, and the scratch tensors (y, dy, uv, ...) as the individual calls leave
 * them.  `side_stream` is reserved and may be NULL: every kernel is enqueued on `stream` (running the blur backward
 * beside the weight-gradient GEMM on a second stream was measured slower than the sequence).  Set struct_bytes =
 * sizeof(neucam_step_args). */
#define NEUCAM_STEP_TO_HIDDEN_WGRAD 1 /* clear, taps, forward, loss, backward chain, hidden-layer weight gradients */
#define NEUCAM_STEP_TAILS 2           /* blur-MLP backward, head weight gradient (a whole-step call runs the blur
                                         backward ahead of the weight-gradient GEMM instead: same results) */
typedef struct neucam_step_args {
  uint32_t struct_bytes;
  int phases;                   /* 0 or 3 = the whole step; NEUCAM_STEP_* bits select a part, so that a data-parallel
                                   caller can start reducing the (large) atlas gradients while the tails still run */
  int B, K, N, H, W;            /* pixels, taps per pixel (9 | 1), frames, frame height / width */
  int out_dim;                  /* atlas output channels (3) */
  int tm_width;                 /* TonemapNet hidden width (128) */
  float offset_size, fixed_value, ue_weight;
  int64_t global_count;         /* 3 * pixels of the GLOBAL batch (mean normaliser) */
  /* step inputs */
  const float* coords;          /* [B,3] (t,y,x) */
  const int64_t* coord_idx;     /* [B] */
  const float* img;             /* [B,3] */
  const float* gamma_w;         /* [B] or NULL (no random-gamma augmentation) */
  const float* gamma_dev;       /* device scalar or NULL */
  const float* mask_w;          /* [B] or NULL */
  /* parameters (fp32 masters) and their tensor-core operand copies */
  const float* focus;           /* [N] (NULL when K == 1) */
  const float* blur_params[8];  /* neucam_blur_fwd order (unused when K == 1) */
  const float *w0, *b0, *b_hid, *w4, *b4;
  const void *w_hid16, *wT_hid16, *w4T16;
  const float* exps;            /* [N] */
  const float* tm_params[12];
  /* gradients (accumulated into) */
  float* g_focus;
  float* g_blur[8];
  float *g_w0, *g_b0, *g_w[3], *g_b[3], *g_w4, *g_b4;
  float* g_exps;
  float* g_tm[12];
  void* zero_ptr;
  int64_t zero_bytes;
  /* scratch owned by the caller */
  float *uv, *kout, *hid, *y, *dy, *dkw, *duv;    /* [K,B,2] [27,B] [192,B] [K,B,3] [K,B,3] [K,B] [K,B,2] */
  void *act16, *phase16, *dz16;
  float* sums;                  /* [4] */
  uint32_t* absmax_bits;        /* [1] */
  float* scale;                 /* [1] loss scale (written) */
  void* ws;                     /* reduction workspace for the main stream: >= max over the NEUCAM_WS_* kinds used */
  int64_t ws_bytes;
  void* ws_side;                /* reduction workspace of the blur backward: NEUCAM_WS_BLUR_BWD */
  int64_t ws_side_bytes;
} neucam_step_args;
int neucam_step_fwd_bwd(const neucam_step_args* args, void* stream, void* side_stream);

#ifdef __cplusplus
}
#endif
#endif /* NEUCAM_B200_H_ */
