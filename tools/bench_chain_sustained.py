"""Does a kernel's duration depend on how long the GPU has been busy?  Times the forward chain in bursts of
increasing length (CUDA events around each burst) and samples the SM clock with NVML while a long burst runs."""
import math, os, sys, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neucam_b200 import ops
import pynvml

def main():
    Q = 382725
    dev = "cuda"
    g = torch.Generator().manual_seed(0)
    w0 = ((torch.rand(512, 2, generator=g) * 2 - 1) * 0.5).to(dev)
    b0 = ((torch.rand(512, generator=g) * 2 - 1) * 0.7).to(dev)
    wh = ((torch.rand(3, 512, 512, generator=g) * 2 - 1) * math.sqrt(6 / 512) / 30).to(dev)
    bh = ((torch.rand(3, 512, generator=g) * 2 - 1) / math.sqrt(512)).to(dev)
    w4 = ((torch.rand(3, 512, generator=g) * 2 - 1) * math.sqrt(6 / 512) / 30).to(dev)
    b4 = torch.zeros(3, device=dev)
    w16 = wh.half().reshape(1536, 512).contiguous()
    uv = (torch.rand(Q, 2, device=dev) * 2 - 1)
    act, cos, dz = ops.alloc_chain_workspace(Q, dev)
    y = torch.empty(Q, 3, device=dev)
    fn = lambda: ops.sine_mlp_fwd(uv, w0, b0, w16, bh, w4, b4, y=y, act16=act, phase16=cos)
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    for n in (1, 3, 10, 30, 100, 300, 1000):
        time.sleep(0.5)
        samples = []
        stop = threading.Event()
        def poll():
            while not stop.is_set():
                samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
                time.sleep(0.002)
        t = threading.Thread(target=poll); t.start()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        stop.set(); t.join()
        ms = e0.elapsed_time(e1) / n
        clk = sorted(s[0] for s in samples); pw = max(s[1] for s in samples)
        print(f"burst of {n:5d} launches: {ms:.3f} ms/launch  {Q * 2 * 788992 / ms / 1e9:7.1f} TFLOP/s   SM clock min/median/max "
              f"{clk[0]}/{clk[len(clk)//2]}/{clk[-1]} MHz over {len(clk)} samples, power max {pw:.0f} W")

if __name__ == "__main__":
    main()
