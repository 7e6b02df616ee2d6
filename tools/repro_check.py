"""Run the same tc-engine step twice on one seeded batch and report what differs (forward outputs bit for bit, gradients
by relative L2), then the per-parameter smoke numbers.  usage: python tools/repro_check.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "li-vae_b200")]
import numpy as np, torch
import livae
from livae.train import rvae_step_loss
from oracle import rvae as O

P, L = 128, 2
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
seed = 5
torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
params = O.make_params(O.rvae_param_shapes(P, L), seed=seed, stn_head_std=0.5)
x, xr, ang = O.make_lattice_batch(B, P, seed=seed + 1)
eps = torch.from_numpy(np.random.default_rng(seed + 2).standard_normal((B, L))).float()
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))

def run():
    m = livae.RVAE(L, 1, P); m.load_state_dict(params); m.cuda()
    orig = torch.randn_like
    torch.randn_like = lambda t, **k: eps.to(device=t.device, dtype=t.dtype).reshape(t.shape)
    try:
        out = rvae_step_loss(m, livae.RVAELoss(beta=10.0, gamma=10.0), x.cuda(), xr.cuda(), ang.cuda(), 0.2)
    finally:
        torch.randn_like = orig
    out[0].backward(); torch.cuda.synchronize()
    o = {k: v.detach().cpu() for k, v in zip(("rotated_recon", "recon", "theta", "mu", "logvar"), out[5])}
    o.update(loss=out[0].detach().cpu(), recon_loss=out[1].detach().cpu(), kld=out[2].detach().cpu(), cycle=out[3].detach().cpu(), canonical=out[4].detach().cpu())
    return o, {k: p.grad.detach().cpu() for k, p in m.named_parameters()}

livae.set_engine("tc")
o1, g1 = run(); o2, g2 = run()
print("forward bit-identical:", {k: bool(torch.equal(o1[k], o2[k])) for k in o1})
print("gradient run-to-run rel L2:")
for k in g1:
    print(f"  {k:46s} {rel(g1[k], g2[k]):.2e}")
want, wg = O.rvae_full_step(params, x, xr, ang, eps)
bf16 = lambda t: t + (t.to(torch.bfloat16).float() - t).detach()
_, qg = O.rvae_full_step(params, x, xr, ang, eps, quant=bf16)
print(f"smoke table at B={B}: ELBO rel {abs(float(o1['loss']) - float(want['loss'])) / float(want['loss']):.2e}")
for k in g1:
    print(f"  {k:46s} tc {rel(g1[k], wg[k]):.2e}   floor {rel(qg[k], wg[k]):.2e}")
