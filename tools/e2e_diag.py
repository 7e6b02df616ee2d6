"""Where the e2e arm's time goes: H2D bandwidth of a pinned batch, device-resident steps with a per-step
loss read-back, and the prefetched host-batch loop (bench.py's e2e arm)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "li-vae_b200")]
import torch
import livae
from livae import optim
from livae.train import train_rvae_step, DevicePrefetcher

B = int(os.environ.get("MB", "2048"))
dev = torch.device("cuda")
torch.manual_seed(0)
model = livae.RVAE(latent_dim=2, in_channels=1, patch_size=128).to(dev)
crit = livae.RVAELoss(beta=10.0, gamma=10.0)
opt = optim.FlatAdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
dbat = [(torch.rand(B, 1, 128, 128, device=dev), torch.rand(B, 1, 128, 128, device=dev), torch.rand(B, device=dev)) for _ in range(3)]
hbat = [tuple(t.cpu().pin_memory() for t in b) for b in dbat]
step = lambda b: train_rvae_step(model, opt, crit, b, dev, 0.2, 20.0, None)
for i in range(4):
    step(dbat[i % 3])
torch.cuda.synchronize()

def ev():
    return torch.cuda.Event(enable_timing=True)

# 1. raw H2D
buf = torch.empty_like(dbat[0][0]); e0, e1 = ev(), ev(); e0.record()
for _ in range(5):
    buf.copy_(hbat[0][0], non_blocking=True)
e1.record(); torch.cuda.synchronize()
print(f"H2D pinned: {5 * buf.numel() * 4 / e0.elapsed_time(e1) / 1e6:.1f} GB/s")
N = 8
for name, feed, read in (("device batches, no read-back", lambda: (dbat[i % 3] for i in range(N)), False),
                         ("device batches, loss.item() every step", lambda: (dbat[i % 3] for i in range(N)), True),
                         ("host batches via DevicePrefetcher, no read-back", lambda: DevicePrefetcher((hbat[i % 3] for i in range(N)), dev), False),
                         ("host batches via DevicePrefetcher, loss.item() every step", lambda: DevicePrefetcher((hbat[i % 3] for i in range(N)), dev), True),
                         ("host batches, plain .to(device) in the step, loss.item()", lambda: (hbat[i % 3] for i in range(N)), True)):
    torch.cuda.synchronize(); e0, e1 = ev(), ev(); t0 = time.perf_counter(); e0.record()
    for b in feed():
        out = step(b)
        if read:
            out[1].item()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / N:.2f} ms/step (wall {1e3 * (time.perf_counter() - t0) / N:.2f})")

# host-side issue time per iteration of the prefetched loop without read-back (is the host ever blocked?)
torch.cuda.synchronize()
ts = []
t0 = time.perf_counter()
for b in DevicePrefetcher((hbat[i % 3] for i in range(12)), dev):
    step(b)
    ts.append(time.perf_counter() - t0)
torch.cuda.synchronize()
tend = time.perf_counter() - t0
print("host issue times (ms):", " ".join(f"{1e3 * (b - a):.1f}" for a, b in zip([0.0] + ts[:-1], ts)), f"| drained at {1e3 * tend:.1f} ms")
