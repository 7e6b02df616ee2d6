"""CPU: the bf16-operand floor of the FULL step's gradients as a function of the batch size (C3 shapes: P = 128, L = 2).

floor(B) = relative L2 distance between oracle/rvae.py in fp32 and the same oracle with every GEMM operand rounded to
bf16 (exact arithmetic otherwise) -- what the best possible engine with bf16 GEMM inputs shows against the reference
(DESIGN 4.1).  Rounding errors of different patches are independent, so the decoder's floor falls like 1/sqrt(B); the
tensor-core engine was measured at 0.8 - 0.9 of it at B = 64 and B = 256 (profiles/r02_final_smoke.txt, r02z_smoke.txt).
Output committed as profiles/r02z_floor_vs_batch_cpu.txt.

    python tools/floor_vs_batch_cpu.py [B ...]          (default 16 64 256 1024; B = 1024 takes a few minutes)
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import rvae as O  # noqa: E402

P, L = 128, 2


def main():
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    sizes = [int(a) for a in sys.argv[1:]] or [16, 64, 256, 1024]
    bf16 = lambda t: t + (t.to(torch.bfloat16).float() - t).detach()
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
    params = O.make_params(O.rvae_param_shapes(P, L), seed=1234, stn_head_std=0.5)
    groups = (("decoder convs d1-d4", lambda k: "deconv_layers" in k), ("decoder.fc", lambda k: k.startswith("decoder.fc")),
              ("latent heads", lambda k: "fc_mu" in k or "fc_logvar" in k), ("encoder convs", lambda k: "encoder.conv" in k),
              ("STN", lambda k: "rotation_stn" in k))
    print(f"{'B':>6s} " + " ".join(f"{g[0]:>22s}" for g in groups) + "   d1.weight   ELBO(rel)   seconds")
    for B in sizes:
        t0 = time.perf_counter()
        x, xr, ang = O.make_lattice_batch(B, P, seed=2024)
        eps = torch.from_numpy(np.random.default_rng(99).standard_normal((B, L))).float()
        o32, g32 = O.rvae_full_step(params, x, xr, ang, eps)
        obf, gbf = O.rvae_full_step(params, x, xr, ang, eps, quant=bf16)
        fl = {k: rel(gbf[k], g32[k]) for k in g32}
        cols = [max(v for k, v in fl.items() if sel(k)) for _, sel in groups]
        print(f"{B:6d} " + " ".join(f"{c:22.2e}" for c in cols) + f"   {fl['decoder.deconv_layers.2.weight']:.2e}"
              f"   {abs(float(obf['loss']) - float(o32['loss'])) / abs(float(o32['loss'])):.1e}   {time.perf_counter() - t0:6.1f}",
              flush=True)


if __name__ == "__main__":
    main()
