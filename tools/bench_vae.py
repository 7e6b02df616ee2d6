"""C2 (BASELINE.json configs[1]): plain VAE, synthetic 64x64 lattice patches, batch 4096, one B200 -- train step
(VAE.forward + VAELoss + backward + clip 5.0 + Adam), patches/s with CUDA events.  Secondary number: bench.py's
headline stays C3.  usage: python tools/bench_vae.py [engine] [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "li-vae_b200")]
import json
import torch
import livae
from livae import optim, ops
from oracle import rvae as O

engine = sys.argv[1] if len(sys.argv) > 1 else "tc"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
P, L = 64, 16
dev = torch.device("cuda")
livae.set_engine(engine)
torch.manual_seed(1234)
m = livae.VAE(latent_dim=L, in_channels=1, patch_size=P).to(dev)
crit = livae.VAELoss(beta=1.0)
opt = optim.FlatAdamW(m.parameters(), lr=1e-3, weight_decay=0.0)
g = torch.Generator(device="cpu").manual_seed(7)
batches = [torch.rand(B, 1, P, P, generator=g).to(dev) for _ in range(4)]     # 4 x 67 MB, cycled


def step(x):
    opt.zero_grad(set_to_none=False)
    recon, mu, logvar = m(x)
    loss, _, _ = crit(recon, x, mu, logvar)
    loss.backward()
    opt.sync_grads()
    ops.l2norm_clip_(opt.flat_grad, 5.0, apply=True)
    opt.step()
    return loss


for i in range(4):
    l0 = step(batches[i % 4])
torch.cuda.synchronize()
N = 10 if engine == "tc" else 3
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(N):
    l1 = step(batches[i % 4])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / N
print(json.dumps({"workload": "C2: VAE P=64 L=16, train step incl. clip 5.0 + Adam", "engine": engine, "batch": B,
                  "ms_per_step": ms, "patches_per_s": B / ms * 1e3, "loss_first": float(l0), "loss_last": float(l1),
                  "step_roofline_ms_fp32_boundaries": 2.62e6 * B / 6551e9 * 1e3,
                  "frac_of_roofline": (2.62e6 * B / 6551e9 * 1e3) / ms}))
if os.environ.get("PROFILE"):
    sys.path.insert(0, ROOT)
    import bench
    fam, _ = bench.profile_families(step, batches, nsteps=2)
    for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])[:25]:
        print(f"{f['ms'] / 2:8.3f} ms {f['calls'] / 2:4.0f}x  {k}")
