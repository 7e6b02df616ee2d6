"""Like-for-like GPU baseline: the reference's rVAE FULL step on stock ATen/cuDNN on this GPU,
fp32 (`--no-amp`) and under autocast fp16 (the reference's default CUDA mode) / bf16, same shapes as bench.py.
Usage: python tools/gpu_baseline.py [B]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "li-vae_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
from oracle import aten_step as A, rvae as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
P, L = 128, 2
dev = torch.device("cuda")
params = O.make_params(O.rvae_param_shapes(P, L), seed=1234, stn_head_std=0.5)
g = torch.Generator(device="cpu").manual_seed(0)
batches = []
for i in range(3):
    x = torch.rand((B, 1, P, P), generator=g).to(dev)
    ang = (torch.rand(B, generator=g) * 2 * np.pi).to(dev)
    xr = A.rot_sample_aten(x, torch.cos(ang), torch.sin(ang))
    batches.append((x, xr, ang))
eps = [torch.randn((B, L), generator=g).to(dev) for _ in range(3)]
res = {}
for name, amp in (("fp32", None), ("amp_fp16", torch.float16), ("amp_bf16", torch.bfloat16)):
    for tf32 in ((False, True) if amp is None else (False,)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True
        key = name + ("_tf32" if tf32 else "")
        try:
            rate, ms = A.time_gpu_baseline(params, batches, eps, amp)
            res[key] = {"patches_per_s": rate, "ms_per_step": ms, "batch": B,
                        "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
        except torch.cuda.OutOfMemoryError:
            res[key] = {"oom": True, "batch": B}
            torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
        print(key, res[key], flush=True)
print(json.dumps(res))

# where the stock-ATen step spends its time: one profiled fp32 step and one autocast-fp16 step
from torch.profiler import ProfilerActivity, profile
for name, amp in (("fp32", None), ("amp_fp16", torch.float16)):
    t = A.AtenTrainer(params, dev, amp=amp)
    x, xr, ang = batches[0]
    for _ in range(2):
        t.step(x, xr, ang, eps[0])
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        t.step(x, xr, ang, eps[0])
        torch.cuda.synchronize()
    print(f"---- top CUDA kernels of one {name} ATen step (B={B})")
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
    del t
    torch.cuda.empty_cache()
