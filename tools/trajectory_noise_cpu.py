"""CPU: how far do two runs of the graph-vs-eager tests' trainings drift apart when their gradients differ by
atomic-order noise only?  (tests/test_gpu_step.py: test_cuda_graph_step_equals_eager_step -- 3 steps -- and
test_train_rvae_one_epoch_with_graphs -- 8 steps incl. a ragged batch; P = 32, B = 16, AdamW lr 1e-3, clip 20.)

The oracle (oracle/rvae.py) stands in for the engines: exact fp32, and with every GEMM operand rounded to bf16 like the
tcgen05 engine.  One run is the yardstick; the others multiply every gradient element by 1 + rel * N(0, 1) each step
(rel = 1e-6 / 1e-5: the measured run-to-run jitter of the backward pass's remaining fp32 atomics).  Printed per
tensor, worst over six noise seeds: mean |difference| / (lr * steps), the fraction of elements beyond a quarter of that
travel, the largest element; and how far the logged losses move.  Output committed as profiles/r02z_trajectory_noise_cpu.txt; it is where the tests' bound
(mean <= 10 % of the travel) comes from, next to the GPU observation of 0.34 %.

    python tools/trajectory_noise_cpu.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import rvae as O  # noqa: E402

P, L, B, LR = 32, 2, 16, 1e-3
bf16 = lambda t: t + (t.to(torch.bfloat16).float() - t).detach()


def schedule(which):
    if which == "step":                                  # three different full batches, one step each
        seed = 77
        return seed, [O.make_lattice_batch(B, P, seed=seed + 1 + k) for k in range(3)]
    seed = 91                                            # two epochs over three full batches and a ragged one
    full = [O.make_lattice_batch(B, P, seed=seed + 1 + k) for k in range(3)]
    return seed, (full + [tuple(t[:B // 2].contiguous() for t in full[0])]) * 2


def train(which, noise_seed, rel, quant):
    seed, batches = schedule(which)
    params = O.make_params(O.rvae_param_shapes(P, L), seed=seed, stn_head_std=0.5)
    eps = torch.from_numpy(np.random.default_rng(seed).standard_normal((B, L))).float()
    m = {k: torch.zeros_like(v) for k, v in params.items()}
    v = {k: torch.zeros_like(v_) for k, v_ in params.items()}
    g = torch.Generator().manual_seed(noise_seed)
    logged = []
    for t, (x, xr, ang) in enumerate(batches, 1):
        kw = dict(beta=10.0, gamma=10.0, canonical_weight=0.2)
        if quant is not None:
            kw["quant"] = quant
        outs, grads = O.rvae_full_step(params, x, xr, ang, eps[:x.shape[0]], **kw)
        logged.append([float(outs[k]) for k in ("loss", "recon_loss", "kld", "cycle")])
        if rel > 0:
            grads = {k: gg * (1 + rel * torch.randn(gg.shape, generator=g)) for k, gg in grads.items()}
        tot = torch.sqrt(sum((gg.double() ** 2).sum() for gg in grads.values())).float()
        coef = torch.clamp(20.0 / (tot + 1e-6), max=1.0)
        for k in params:                                 # AdamW as scripts/train_rvae.py:157-159 configures it
            gg = grads[k] * coef
            params[k].mul_(1 - LR * 1e-5)
            m[k].lerp_(gg, 0.1)
            v[k].mul_(0.999).addcmul_(gg, gg, value=0.001)
            params[k].addcdiv_(m[k] / (1 - 0.9 ** t), (v[k] / (1 - 0.999 ** t)).sqrt_().add_(1e-8), value=-LR)
    logged = np.asarray(logged)
    if which == "epoch":                                 # the epoch test compares epoch AVERAGES (two epochs of four batches)
        logged = logged.reshape(2, 4, -1).mean(1)
    return params, LR * len(batches), logged


def main():
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    for which in ("step", "epoch"):
        for quant, qn in ((None, "fp32 oracle"), (bf16, "bf16-operand oracle")):
            base, travel, base_log = train(which, 0, 0.0, quant)
            for rel in (1e-6, 1e-5):
                worst = {k: (0.0, 0.0, 0.0) for k in base}
                wlog = np.zeros(4)
                for s in range(1, 7):
                    p, _, log = train(which, s, rel, quant)
                    wlog = np.maximum(wlog, (np.abs(log - base_log) / np.abs(base_log)).max(0))
                    for k in base:
                        d = (p[k] - base[k]).abs()
                        st = (float(d.mean()) / travel, float((d > 0.25 * travel).float().mean()), float(d.max()) / travel)
                        worst[k] = tuple(max(a, b) for a, b in zip(worst[k], st))
                top = sorted(base, key=lambda k: -worst[k][0])[:4]
                print(f"{which:5s} {qn:20s} noise {rel:.0e}: overall worst mean/travel {max(w[0] for w in worst.values()):.4f}, "
                      f"fraction beyond travel/4 {max(w[1] for w in worst.values()):.4f}, max/travel {max(w[2] for w in worst.values()):.2f}; "
                      + "; ".join(f"{k} {worst[k][0]:.4f}" for k in top)
                      + "; logged loss / recon / kld / cycle differ by at most (relative) "
                      + " / ".join(f"{w:.1e}" for w in wlog))


if __name__ == "__main__":
    main()
