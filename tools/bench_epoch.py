"""The drop-in trainer itself: livae.train.train_rvae_one_epoch (reference train.py:286-445) over a list of pinned host
batches (what a DataLoader with pin_memory yields), including the per-step metric block (PSNR, SSIM, canonical
PSNR/SSIM, latent / rotation statistics) and the epoch-end metric read-back.
usage: python tools/bench_epoch.py [nbatches] [host|device]
  host   (default): batches come from pinned host memory (H2D inside the loop)
  device: the C5 flow -- 16 synthetic 4096x4096 images resident in HBM, float sites, every batch produced by
          livae.data.DevicePatchLoader (ROI gather + default_transform + paired rotation + min-max on the device,
          draws from Python random in the reference's order), no host->device pixel traffic at all"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "li-vae_b200")]
import numpy as np
import torch
import livae
from livae import optim, ops

NB = int(sys.argv[1]) if len(sys.argv) > 1 else 12
MODE = sys.argv[2] if len(sys.argv) > 2 else "host"
B, P = 2048, 128
dev = torch.device("cuda")
torch.manual_seed(1234)
m = livae.RVAE(latent_dim=2, in_channels=1, patch_size=P).to(dev)
crit = livae.RVAELoss(beta=10.0, gamma=10.0)
opt = optim.FlatAdamW(m.parameters(), lr=1e-3, weight_decay=1e-5)
if MODE == "device":
    import random
    from bench import synth_haadf
    from livae.data import DevicePatchLoader, DevicePatchSource, default_transform
    n_img, hw = 16, 4096
    imgs = torch.stack([synth_haadf(hw, 12.0, 3.75 * k, 1000 + k, dev) for k in range(n_img)]).contiguous()
    rng = np.random.default_rng(5)
    per_img = (NB + 3) * B // n_img + 1
    coords = [rng.uniform(96, hw - 96, size=(per_img, 2)) for _ in range(n_img)]
    src = DevicePatchSource(imgs, coords, P, 32, transform=default_transform, device=dev)
    random.seed(11)
    warm = DevicePatchLoader(src, B, mode="paired", seed=1, indices=np.arange(3 * B))
    loader = DevicePatchLoader(src, B, mode="paired", seed=2, indices=np.arange(3 * B, (NB + 3) * B))
    # the data side alone
    for _b in DevicePatchLoader(src, B, mode="paired", seed=3, indices=np.arange(2 * B)):     # warm-up (allocations)
        pass
    torch.cuda.synchronize(); t0 = time.perf_counter()
    nb = 0
    for _b in DevicePatchLoader(src, B, mode="paired", seed=3, indices=np.arange(8 * B) % (3 * B)):
        nb += 1
    torch.cuda.synchronize()
    data_rate = nb * B / (time.perf_counter() - t0)
    what = ("train_rvae_one_epoch fed by livae.data.DevicePatchLoader (16 x 4096^2 images resident in HBM, ROI gather + "
            "default_transform + paired rotation + min-max on the device), metric block included")
else:
    g = torch.Generator(device="cpu").manual_seed(3)
    host = []
    for _ in range(3):
        x = torch.rand(B, 1, P, P, generator=g).to(dev)
        ang = (torch.rand(B, generator=g) * 2 * np.pi).to(dev)
        xr = ops.rot_sample(x, ops.angle_to_cs(ang), 1.0)
        host.append(tuple(t.cpu().pin_memory() for t in (x, xr, ang)))
    loader = [host[i % 3] for i in range(NB)]
    warm = loader[:3]
    data_rate = None
    what = "train_rvae_one_epoch, C3 batches from pinned host memory, metric block included"
log = livae.MetricLogger()
livae.train_rvae_one_epoch(m, warm, opt, crit, log, dev)       # warm-up epoch
torch.cuda.synchronize()
t0 = time.perf_counter()
livae.train_rvae_one_epoch(m, loader, opt, crit, log, dev)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(json.dumps({"workload": what, "data_pipeline_alone_patches_per_s": data_rate,
                  "batches": NB, "ms_per_step": 1e3 * dt / NB, "patches_per_s": NB * B / dt,
                  "metrics": {k: round(v[-1], 5) for k, v in log.metrics.items()}}))
