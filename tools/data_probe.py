"""per-kernel times of one livae.data.DevicePatchSource.paired_batch (B=2048, P=128, padding 32, 4096^2 images)"""
import os, sys, time, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "li-vae_b200")]
import numpy as np
import torch
from bench import synth_haadf
from livae import _lib
from livae.data import DevicePatchSource, default_transform

dev = torch.device("cuda")
B, P, hw, n_img = 2048, 128, 4096, 4
imgs = torch.stack([synth_haadf(hw, 12.0, 3.75 * k, 1000 + k, dev) for k in range(n_img)]).contiguous()
rng = np.random.default_rng(5)
coords = [rng.uniform(96, hw - 96, size=(4 * B, 2)) for _ in range(n_img)]
src = DevicePatchSource(imgs, coords, P, 32, transform=default_transform, device=dev)
random.seed(1)
for k in range(2):
    src.paired_batch(np.arange(k * B, (k + 1) * B))
torch.cuda.synchronize()
_lib.PROFILE = []
t0 = time.perf_counter()
src.paired_batch(np.arange(2 * B, 3 * B))
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
rec, _lib.PROFILE = _lib.PROFILE, None
for name, a, e0, e1 in rec:
    print(f"{name:32s} {e0.elapsed_time(e1):7.3f} ms")
print(f"host side of the call {1e3 * (t1 - t0):.2f} ms, until the device is idle {1e3 * (t2 - t0):.2f} ms")
