"""time of the per-step metric block of train_rvae_one_epoch (reference train.py:399-427): C-ABI calls by name plus
the total step time with and without it"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "li-vae_b200")]
import numpy as np
import torch
import livae
from livae import _lib, optim, ops
from livae.train import train_rvae_step

B, P = 2048, 128
dev = torch.device("cuda")
torch.manual_seed(1234)
m = livae.RVAE(latent_dim=2, in_channels=1, patch_size=P).to(dev)
crit = livae.RVAELoss(beta=10.0, gamma=10.0)
opt = optim.FlatAdamW(m.parameters(), lr=1e-3, weight_decay=1e-5)
g = torch.Generator(device="cpu").manual_seed(3)
x = torch.rand(B, 1, P, P, generator=g).to(dev)
ang = (torch.rand(B, generator=g) * 2 * np.pi).to(dev)
xr = ops.rot_sample(x, ops.angle_to_cs(ang), 1.0)
loader = [(x, xr, ang)] * 6
log = livae.MetricLogger()
livae.train_rvae_one_epoch(m, loader[:2], opt, crit, log, dev)
torch.cuda.synchronize()


def timed(fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


t_epoch = timed(lambda: livae.train_rvae_one_epoch(m, loader, opt, crit, log, dev), 6)
t_step = timed(lambda: [train_rvae_step(m, opt, crit, b, dev, 0.2, 20.0, None) for b in loader], 6)
print(f"train_rvae_one_epoch {t_epoch:.2f} ms/step, bare train_rvae_step {t_step:.2f} ms/step")
_lib.PROFILE = []
livae.train_rvae_one_epoch(m, loader[:1], opt, crit, log, dev)
torch.cuda.synchronize()
rec, _lib.PROFILE = _lib.PROFILE, None
agg = {}
for name, a, e0, e1 in rec:
    if name in ("livae_ssim_box", "livae_rot_sample_fwd", "livae_elbo_fwd", "livae_angle_to_cs"):
        agg.setdefault(name, []).append(e0.elapsed_time(e1))
for k, v in agg.items():
    print(k, [round(t, 3) for t in v])
