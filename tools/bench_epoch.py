"""The drop-in trainer itself: livae.train.train_rvae_one_epoch (reference train.py:286-445) over a list of pinned host
batches (what a DataLoader with pin_memory yields), including the per-step metric block (PSNR, SSIM, canonical
PSNR/SSIM, latent / rotation statistics) and the epoch-end metric read-back.  usage: python tools/bench_epoch.py [nbatches]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "li-vae_b200")]
import numpy as np
import torch
import livae
from livae import optim, ops

NB = int(sys.argv[1]) if len(sys.argv) > 1 else 12
B, P = 2048, 128
dev = torch.device("cuda")
torch.manual_seed(1234)
m = livae.RVAE(latent_dim=2, in_channels=1, patch_size=P).to(dev)
crit = livae.RVAELoss(beta=10.0, gamma=10.0)
opt = optim.FlatAdamW(m.parameters(), lr=1e-3, weight_decay=1e-5)
g = torch.Generator(device="cpu").manual_seed(3)
host = []
for _ in range(3):
    x = torch.rand(B, 1, P, P, generator=g).to(dev)
    ang = (torch.rand(B, generator=g) * 2 * np.pi).to(dev)
    xr = ops.rot_sample(x, ops.angle_to_cs(ang), 1.0)
    host.append(tuple(t.cpu().pin_memory() for t in (x, xr, ang)))
loader = [host[i % 3] for i in range(NB)]
log = livae.MetricLogger()
livae.train_rvae_one_epoch(m, loader[:3], opt, crit, log, dev)       # warm-up epoch
torch.cuda.synchronize()
t0 = time.perf_counter()
livae.train_rvae_one_epoch(m, loader, opt, crit, log, dev)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(json.dumps({"workload": "train_rvae_one_epoch, C3 batches from pinned host memory, metric block included",
                  "batches": NB, "ms_per_step": 1e3 * dt / NB, "patches_per_s": NB * B / dt,
                  "metrics": {k: round(v[-1], 5) for k, v in log.metrics.items()}}))
