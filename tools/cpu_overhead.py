"""Host-side cost of one FULL rVAE train step: with a tiny batch the GPU work is negligible, so the wall time per
step is the Python + autograd + ctypes launch overhead that a large-batch step must hide behind GPU time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "li-vae_b200")]
import torch
import livae
from livae import optim, _lib
from livae.train import train_rvae_step

B = int(os.environ.get("MB", "8"))
dev = torch.device("cuda")
torch.manual_seed(0)
model = livae.RVAE(latent_dim=2, in_channels=1, patch_size=128).to(dev)
crit = livae.RVAELoss(beta=10.0, gamma=10.0)
opt = optim.FlatAdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
x = torch.rand(B, 1, 128, 128, device=dev); xr = torch.rand(B, 1, 128, 128, device=dev); ang = torch.rand(B, device=dev)
for _ in range(5):
    train_rvae_step(model, opt, crit, (x, xr, ang), dev, 0.2, 20.0, None)
torch.cuda.synchronize()
L = _lib.lib()
c0 = L.livae_launch_count()
t0 = time.perf_counter()
N = 20
for _ in range(N):
    train_rvae_step(model, opt, crit, (x, xr, ang), dev, 0.2, 20.0, None)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"batch {B}: host {1e3 * (t1 - t0) / N:.2f} ms/step issue, {1e3 * (t2 - t0) / N:.2f} ms/step incl. drain, "
      f"{(L.livae_launch_count() - c0) / N:.0f} library launches/step")
if len(sys.argv) > 1 and sys.argv[1] == "prof":
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    for _ in range(10):
        train_rvae_step(model, opt, crit, (x, xr, ang), dev, 0.2, 20.0, None)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
