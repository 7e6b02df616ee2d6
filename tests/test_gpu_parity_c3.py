"""GPU parity of the BENCHMARKED configuration (C3: rVAE P=128, L=2, FULL step; tcgen05 engine) against the oracle
at B = 256 -- large enough that per-sample rounding noise averages and the CPU oracle still finishes in seconds.

Three yardsticks, all computed on the same seeded batch:
  oracle32     oracle/rvae.py in fp32 on the CPU: the reference algorithm.
  oracle_bf16  the same oracle with every GEMM operand (activations and weights of conv / linear layers) rounded to
               bf16 and exact arithmetic otherwise: what ANY engine with bf16 GEMM inputs computes at best.  Its
               distance from oracle32 is the operand-rounding FLOOR of a parameter's gradient.  The floor is large
               below the latent bottleneck (1e-1 class for the encoder and the STN, 4e-2..7e-2 for decoder.fc): the
               gradient there is a heavily cancelling sum, and max-pool / ReLU routing decisions flip under 1e-3
               perturbations.  The reference's own autocast modes sit at or above it (profiles/r02_parity_c3_b256.txt).
  aten32       the reference's ops (ATen/cuDNN, F.affine_grid + F.grid_sample) on THIS GPU in fp32.

Bars (north_star): ELBO 1e-3; reconstructions 1e-2 (bf16 GEMM inputs); the decoder's convolution gradients 1e-2 or
1.5x their floor; every other gradient max(1e-2, 3x floor).  The exact engine is held to the reference's own
CPU-vs-GPU fp32 difference (3x |aten32 - oracle32|, at least 1e-3) and to 1e-3 on the decoder outright."""
import os

import numpy as np
import pytest
import torch

from oracle import aten_step as A
from oracle import rvae as O
from tests.test_gpu_step import FixedEps
from tests.util import rel_l2

pytestmark = pytest.mark.gpu
P, L, B = 128, 2, 256


def _ste(dt):
    return lambda t: t + (t.to(dt).float() - t).detach()


@pytest.fixture(scope="module")
def case():
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    params = O.make_params(O.rvae_param_shapes(P, L), seed=1234, stn_head_std=0.5)
    x, xr, ang = O.make_lattice_batch(B, P, seed=2024)
    eps = torch.from_numpy(np.random.default_rng(99).standard_normal((B, L))).float()
    o32, g32 = O.rvae_full_step(params, x, xr, ang, eps)
    obf, gbf = O.rvae_full_step(params, x, xr, ang, eps, quant=_ste(torch.bfloat16))
    floor = {k: rel_l2(gbf[k], g32[k]) for k in g32}
    t = A.AtenTrainer(params, "cuda", amp=None)
    t.forward_backward(x.cuda(), xr.cuda(), ang.cuda(), eps.cuda())
    gat = {k: v.detach().cpu() for k, v in t.grads().items()}
    return dict(params=params, x=x, xr=xr, ang=ang, eps=eps, o32=o32, g32=g32, obf=obf, gbf=gbf, floor=floor, gat=gat)


def _run(engine, c):
    import livae
    from livae.train import rvae_step_loss
    livae.set_engine(engine)
    try:
        m = livae.RVAE(L, 1, P)
        m.load_state_dict(c["params"])
        m.cuda()
        crit = livae.RVAELoss(beta=10.0, gamma=10.0)
        with FixedEps(c["eps"]):
            loss, rl, kl, cyc, can, outs = rvae_step_loss(m, crit, c["x"].cuda(), c["xr"].cuda(), c["ang"].cuda(), 0.2)
        loss.backward()
        torch.cuda.synchronize()
        g = {k: p.grad.detach().cpu() for k, p in m.named_parameters()}
        o = dict(loss=float(loss), recon_loss=float(rl), kld=float(kl), cycle=float(cyc), canonical=float(can),
                 rotated_recon=outs[0].detach().cpu(), recon=outs[1].detach().cpu(), theta=outs[2].detach().cpu(),
                 mu=outs[3].detach().cpu(), logvar=outs[4].detach().cpu())
    finally:
        livae.set_engine("tc")
    return o, g


def test_operand_rounding_floor_is_what_the_docstring_says(case):
    """the premise of the bounds below, checked rather than asserted in prose"""
    f = case["floor"]
    assert max(v for k, v in f.items() if "deconv_layers" in k) < 3e-2
    assert max(f["decoder.fc.weight"], f["decoder.fc.bias"]) > 1e-2
    assert max(v for k, v in f.items() if k.startswith("encoder.")) > 5e-2
    # ... while the forward quantities survive bf16 operands easily
    o32, obf = case["o32"], case["obf"]
    assert abs(float(obf["loss"]) - float(o32["loss"])) < 1e-4 * abs(float(o32["loss"]))
    assert rel_l2(obf["rotated_recon"], o32["rotated_recon"]) < 2e-3


def test_tensor_core_engine_at_benchmark_shapes(case):
    o32, g32, floor = case["o32"], case["g32"], case["floor"]
    o, g = _run("tc", case)
    # per-batch ELBO and its terms (north_star: 1e-3)
    assert abs(o["loss"] - float(o32["loss"])) <= 1e-3 * abs(float(o32["loss"]))
    assert abs(o["recon_loss"] - float(o32["recon_loss"])) <= 1e-3 * float(o32["recon_loss"])
    assert abs(o["kld"] - float(o32["kld"])) <= 1e-2 * float(o32["kld"]) + 1e-8
    assert abs(o["cycle"] - float(o32["cycle"])) <= 2e-3
    assert abs(o["canonical"] - float(o32["canonical"])) <= 1e-2 * float(o32["canonical"])
    # reconstructions (1e-2 where bf16 GEMM inputs are used)
    assert rel_l2(o["rotated_recon"], o32["rotated_recon"]) <= 1e-2
    assert rel_l2(o["recon"], o32["recon"]) <= 1e-2
    dth = (o["theta"].reshape(-1) - o32["theta"].reshape(-1)).abs()
    dth = torch.minimum(dth, 2 * np.pi - dth)
    assert float(dth.median()) <= 1e-2
    # gradients
    bad = []
    for k in g32:
        e = rel_l2(g[k], g32[k])
        tol = max(1e-2, 1.5 * floor[k]) if "deconv_layers" in k else max(1e-2, 3.0 * floor[k])
        if e > tol:
            bad.append((k, e, tol))
    assert not bad, bad
    # the loss-weighted picture: error of the WHOLE gradient vector relative to its norm
    num = sum(float((g[k].double() - g32[k].double()).pow(2).sum()) for k in g32)
    den = sum(float(g32[k].double().pow(2).sum()) for k in g32)
    assert (num / den) ** 0.5 <= 3.0 * (sum(float((case["gbf"][k].double() - g32[k].double()).pow(2).sum()) for k in g32) / den) ** 0.5


def test_tensor_core_engine_is_no_further_from_bf16_oracle_than_the_floor(case):
    """two bf16-operand evaluations (this engine, the operand-rounded oracle) differ from each other by no more than
    each differs from fp32: the engine adds no error of its own class on top of operand rounding"""
    g32, gbf, floor = case["g32"], case["gbf"], case["floor"]
    _, g = _run("tc", case)
    for k in g32:
        assert rel_l2(g[k], gbf[k]) <= max(1e-2, 3.0 * floor[k]), k


def test_exact_engine_at_benchmark_shapes(case):
    o32, g32, gat = case["o32"], case["g32"], case["gat"]
    o, g = _run("f32", case)
    assert abs(o["loss"] - float(o32["loss"])) <= 1e-5 * abs(float(o32["loss"]))
    assert rel_l2(o["rotated_recon"], o32["rotated_recon"]) <= 1e-4
    assert rel_l2(o["mu"], o32["mu"]) <= 2e-3 and rel_l2(o["logvar"], o32["logvar"]) <= 2e-3
    for k in g32:
        e = rel_l2(g[k], g32[k])
        if k.startswith("decoder."):
            assert e <= 1e-3, (k, e)
        else:
            # below the STN the reference's own fp32 result moves by this much between the CPU and this GPU
            ref_move = rel_l2(gat[k], g32[k])
            assert e <= max(1e-3, 3.0 * ref_move), (k, e, ref_move)


def test_forward_pass_is_bit_reproducible_and_gradients_only_jitter_by_summation_order(case):
    """The reference's CPU path is bit-reproducible.  Here every reduction that is split over CTAs in the
    FORWARD pass (split-K Linear layers: the STN's fc1, the latent heads) and in the tcgen05 weight gradients / bias
    column sums is finished in a fixed order through the scratch buffer (livae_set_scratch), so two runs of the same
    step give identical theta, mu, reconstructions and loss -- hence identical ReLU / max-pool decisions.  The
    backward pass still has fp32 atomics (rot_sample's shared-memory scatter, the thin 1-channel layers, d4,
    decoder.fc, the STN tail), so gradients differ run to run by summation order only: 1e-6 class, no longer
    amplified through flipped activations (round 1 measured 2e-3..5e-3 for the STN here)."""
    o1, g1 = _run("tc", case)
    o2, g2 = _run("tc", case)
    for k in ("theta", "mu", "logvar", "recon", "rotated_recon"):
        assert torch.equal(o1[k], o2[k]), k
    assert o1["loss"] == o2["loss"] and o1["kld"] == o2["kld"] and o1["cycle"] == o2["cycle"]
    for k in g1:
        # measured: <= 5e-7 (B = 64) .. 2e-5 (B = 256, STN conv1 bias) everywhere except the encoder's convolutions and heads, whose gradients are cancelling sums
        # thousands of times smaller than their terms (|g| ~ 1e-2 .. 2e-1 against 2e+1 for the decoder): 2e-4 / 6.5e-5
        # The STN's gradients are the same kind of sum (over samples, of per-sample angle gradients that differ run to run
        # at 1e-7 through rot_sample's atomics): mostly bit-identical, 1.1e-4 seen once in three runs at B = 256.
        loose = k.startswith("encoder.")
        assert rel_l2(g1[k], g2[k]) <= (2e-3 if loose else 1e-4), (k, rel_l2(g1[k], g2[k]))
