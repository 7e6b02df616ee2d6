"""theta / mu error of the tc engine against the oracle over several seeds, thin layers on tcgen05 vs SIMT"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "li-vae_b200")]
import numpy as np, torch
import livae
from livae import _lib
from oracle import rvae as O
L_ = _lib.lib()
P, L, B = 128, 2, 8
livae.set_engine("tc")
for seed in (1234, 1, 2, 3, 4, 5):
    params = O.make_params(O.rvae_param_shapes(P, L), seed=seed, stn_head_std=0.5)
    x, xr, ang = O.make_lattice_batch(B, P, seed=seed + 1)
    m0 = livae.RVAE(L, 1, P); m0.load_state_dict(params)
    th_w = None
    m = m0.cuda()
    res = []
    for mode in (0, 1):
        L_.livae_thin_set_tc(mode)
        with torch.no_grad():
            mu, lv, th = m.encoder(x.cuda())
        res.append((mu.cpu(), th.cpu()))
    if th_w is None:
        livae.set_engine("f32")
        with torch.no_grad():
            mu_w, lv_w, th_w = [t.cpu() for t in m.encoder(x.cuda())]
        livae.set_engine("tc")
    for name, (mu, th) in zip(("simt", "tc  "), res):
        e = (th - th_w).abs().flatten()
        em = (mu - mu_w).abs().flatten()
        print(f"seed {seed} thin={name} theta err med {e.median():.4f} max {e.max():.4f} | mu err med {em.median():.5f} max {em.max():.5f}", flush=True)
