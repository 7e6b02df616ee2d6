"""Parity table at the benchmark's shapes (C3: rVAE P=128, L=2, FULL step) on one GPU.

For one seeded batch of B patches it evaluates the same train-step body (loss + all 30 gradients) with
  oracle32   : oracle/rvae.py on the CPU in fp32 -- the reference algorithm (the yardstick)
  oracle_bf16: the same with every GEMM operand rounded to bf16 (exact arithmetic otherwise) = what ANY engine
               with bf16 GEMM inputs computes at best; |oracle_bf16 - oracle32| is the operand-rounding FLOOR
  aten32/aten_fp16/aten_bf16: stock ATen/cuDNN on this GPU, fp32 and under torch.autocast (the reference's
               own default CUDA mode is autocast fp16, train.py:343-371)
  f32 / tc   : this repo's two engines through the C ABI
and prints, per parameter, the relative L2 distance of each gradient from oracle32 (and of `tc` from
oracle_bf16).  Usage: python tools/parity_c3.py [B] [out.json]
"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "li-vae_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import aten_step as A  # noqa: E402
from oracle import rvae as O  # noqa: E402

P, L = 128, 2


def ste(dt):
    return lambda t: t + (t.to(dt).float() - t).detach()


def rel(a, b):
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def run_engine(engine, params, x, xr, ang, eps):
    import livae
    from livae.train import rvae_step_loss
    livae.set_engine(engine)
    m = livae.RVAE(L, 1, P)
    m.load_state_dict(params)
    m.cuda()
    crit = livae.RVAELoss(beta=10.0, gamma=10.0)
    orig = torch.randn_like
    torch.randn_like = lambda t, **k: eps.to(device=t.device, dtype=t.dtype).reshape(t.shape)
    try:
        loss, rl, kl, cyc, can, outs = rvae_step_loss(m, crit, x.cuda(), xr.cuda(), ang.cuda(), 0.2)
    finally:
        torch.randn_like = orig
    loss.backward()
    torch.cuda.synchronize()
    rotated, recon, theta, mu, logvar = outs
    o = dict(loss=loss, recon_loss=rl, kld=kl, cycle=cyc, canonical=can, rotated_recon=rotated, recon=recon,
             theta=theta, mu=mu, logvar=logvar)
    g = {k: p.grad.detach().cpu() for k, p in m.named_parameters()}
    livae.set_engine("tc")
    return {k: v.detach().cpu() for k, v in o.items()}, g


def run_aten(params, x, xr, ang, eps, amp):
    t = A.AtenTrainer(params, "cuda", amp=amp)
    for _ in range(12):
        loss, aux = t.forward_backward(x.cuda(), xr.cuda(), ang.cuda(), eps.cuda())
        if all(bool(torch.isfinite(g).all()) for g in t.grads().values()):
            break
        # fp16 autocast: the GradScaler starts at 2**16 and overflows; it halves its scale on every such step
        # (the reference skips those optimiser steps the same way, train.py:366-371)
        t.scaler.step(t.opt)
        t.scaler.update()
    o = dict(loss=loss, recon_loss=aux["recon"], kld=aux["kld"], cycle=aux["cycle"], canonical=aux["canonical"],
             rotated_recon=aux["rotated_recon"], recon=aux["canonical_recon"], theta=aux["theta"], mu=aux["mu"],
             logvar=aux["logvar"])
    g = {k: v.detach().float().cpu() for k, v in t.grads().items()}
    return {k: v.detach().float().cpu() for k, v in o.items()}, g


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    out_path = sys.argv[2] if len(sys.argv) > 2 else None
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    params = O.make_params(O.rvae_param_shapes(P, L), seed=1234, stn_head_std=0.5)
    x, xr, ang = O.make_lattice_batch(B, P, seed=2024)
    eps = torch.from_numpy(np.random.default_rng(99).standard_normal((B, L))).float()
    t0 = time.time()
    o32, g32 = O.rvae_full_step(params, x, xr, ang, eps)
    obf, gbf = O.rvae_full_step(params, x, xr, ang, eps, quant=ste(torch.bfloat16))
    print(f"oracle: {time.time() - t0:.1f} s for two CPU steps at B={B}", flush=True)
    runs = {"oracle_bf16": (obf, gbf)}
    for name, amp in (("aten32", None), ("aten_fp16", torch.float16), ("aten_bf16", torch.bfloat16)):
        runs[name] = run_aten(params, x, xr, ang, eps, amp)
    for eng in ("f32", "tc"):
        if eng == "f32" and B > 512:
            continue
        runs[eng] = run_engine(eng, params, x, xr, ang, eps)
    table = {"B": B, "outputs": {}, "grads": {}, "tc_vs_oracle_bf16": {}}
    for name, (o, g) in runs.items():
        table["outputs"][name] = {
            "loss_rel": abs(float(o["loss"]) - float(o32["loss"])) / abs(float(o32["loss"])),
            "recon_loss_rel": abs(float(o["recon_loss"]) - float(o32["recon_loss"])) / abs(float(o32["recon_loss"])),
            "kld_rel": abs(float(o["kld"]) - float(o32["kld"])) / abs(float(o32["kld"])),
            "cycle_abs": abs(float(o["cycle"]) - float(o32["cycle"])),
            "canonical_rel": abs(float(o["canonical"]) - float(o32["canonical"])) / abs(float(o32["canonical"])),
            "rotated_recon": rel(o["rotated_recon"], o32["rotated_recon"]), "recon": rel(o["recon"], o32["recon"]),
            "mu": rel(o["mu"], o32["mu"]), "logvar": rel(o["logvar"], o32["logvar"]),
            "theta_abs_median": float((o["theta"].reshape(-1) - o32["theta"].reshape(-1)).abs().median()),
            "theta_abs_max": float((o["theta"].reshape(-1) - o32["theta"].reshape(-1)).abs().max()),
        }
        table["grads"][name] = {k: rel(g[k], g32[k]) for k in g32}
    gtot = float(torch.sqrt(sum((v.double() ** 2).sum() for v in g32.values())))
    table["grad_norm_total"] = gtot
    table["grad_norms"] = {k: float(v.norm()) for k, v in g32.items()}
    if "tc" in runs:
        table["tc_vs_oracle_bf16"] = {k: rel(runs["tc"][1][k], gbf[k]) for k in g32}
    for eng in ("f32", "tc"):
        if eng in runs:
            table[eng + "_vs_aten32"] = {k: rel(runs[eng][1][k], runs["aten32"][1][k]) for k in g32}
    names = list(runs)
    print("outputs (distance from oracle32):")
    for k in table["outputs"][names[0]]:
        print(f"  {k:18s} " + "  ".join(f"{n}:{table['outputs'][n][k]:.2e}" for n in names))
    print("gradients (relative L2 distance from oracle32)   [last column: tc vs oracle_bf16]")
    for k in g32:
        print(f"  {k:46s} |g|={table['grad_norms'][k]:.2e} " + " ".join(f"{n}:{table['grads'][n][k]:.1e}" for n in names)
              + (f"  | {table['tc_vs_oracle_bf16'][k]:.1e}" if table["tc_vs_oracle_bf16"] else "")
              + "".join(f" {e}~aten32:{table[e + '_vs_aten32'][k]:.1e}" for e in ("f32", "tc") if e + "_vs_aten32" in table))
    if out_path:
        with open(out_path, "w") as f:
            json.dump(table, f, indent=1)


if __name__ == "__main__":
    main()
