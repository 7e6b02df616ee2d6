"""CPU: pin the oracle restatement against vectors produced by the real reference
(tests/golden/make_golden.py).  No GPU, no reference needed at run time."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import patch as OP
from oracle import rot_sample as ORS
from oracle import rvae as O
from tests.golden.make_golden import synth_image
from tests.util import check_grads_against_golden, fp32_noise_floor, load_golden, rel_l2


def test_rot_sample_numpy_matches_reference():
    g = load_golden("rot_sample.npz")
    th = g["thetas"].astype(np.float64)
    # c, s as the float32 values the reference used
    c = np.cos(th.astype(np.float32)).astype(np.float64)
    s = np.sin(th.astype(np.float32)).astype(np.float64)
    for HW in (16, 32):
        rng = np.random.default_rng(7 + HW)
        x = rng.random((len(th), 1, HW, HW)).astype(np.float32)
        go = rng.standard_normal((len(th), 1, HW, HW)).astype(np.float32)
        out = ORS.rot_sample_fwd(x, c, s)
        assert np.abs(out - g[f"out{HW}"]).max() < 2e-5
        gx, gc, gs, _, _ = ORS.rot_sample_bwd(x, c, s, go)
        assert np.abs(gx - g[f"gx{HW}"]).max() < 5e-5
        # at exact multiples of pi/2 every sample lands on a pixel centre, where the
        # coordinate gradient is discontinuous (left/right derivatives differ); only
        # generic angles have a well-defined d/dcos, d/dsin
        gen = [4, 5, 6, 7]
        assert rel_l2(gc[gen], g[f"gc{HW}"][gen]) < 1e-4
        assert rel_l2(gs[gen], g[f"gs{HW}"][gen]) < 1e-4


def test_rot_sample_torch_matches_numpy():
    rng = np.random.default_rng(3)
    x = torch.from_numpy(rng.random((5, 2, 24, 24))).double().requires_grad_(True)
    th = torch.tensor([0.1, 1.0, -2.0, 3.0, 0.77], dtype=torch.float64)
    c = torch.cos(th).requires_grad_(True); s = torch.sin(th).requires_grad_(True)
    go = torch.from_numpy(rng.standard_normal((5, 2, 24, 24))).double()
    out = O.rot_sample_t(x, c, s)
    (out * go).sum().backward()
    o2 = ORS.rot_sample_fwd(x.detach().numpy(), c.detach().numpy(), s.detach().numpy())
    gx, gc, gs, _, _ = ORS.rot_sample_bwd(x.detach().numpy(), c.detach().numpy(),
                                          s.detach().numpy(), go.numpy())
    assert np.abs(out.detach().numpy() - o2).max() < 1e-12
    assert np.abs(x.grad.numpy() - gx).max() < 1e-12
    assert np.abs(c.grad.numpy() - gc).max() < 1e-10
    assert np.abs(s.grad.numpy() - gs).max() < 1e-10


def _rvae_case(tag):
    g = load_golden(f"rvae_step_{tag}.npz")
    P, L, B, seed = int(g["P"]), int(g["L"]), int(g["B"]), int(g["seed"])
    params = O.make_params(O.rvae_param_shapes(P, L), seed=seed, stn_head_std=0.5)
    x, xr, ang = O.make_lattice_batch(B, P, seed=seed + 1)
    eps = torch.from_numpy(np.random.default_rng(seed + 2).standard_normal((B, L))).float()
    return g, params, x, xr, ang, eps


def _check_rvae(tag):
    g, params, x, xr, ang, eps = _rvae_case(tag)
    outs, grads = O.rvae_full_step(params, x, xr, ang, eps, beta=10.0, gamma=10.0,
                                   canonical_weight=0.2)
    assert rel_l2(outs["theta"], g["theta"]) < 1e-5
    assert rel_l2(outs["theta_rot"], g["theta_rot"]) < 1e-5
    assert rel_l2(outs["mu"], g["mu"]) < 1e-4
    assert rel_l2(outs["logvar"], g["logvar"]) < 1e-4
    assert abs(float(outs["loss"]) - float(g["metric/train_loss"])) <= 1e-5 * abs(float(g["metric/train_loss"]))
    assert abs(float(outs["recon_loss"]) - float(g["metric/train_recon_loss"])) <= 1e-5 * float(g["metric/train_recon_loss"])
    assert abs(float(outs["kld"]) - float(g["metric/train_kld_loss"])) <= 1e-4 * float(g["metric/train_kld_loss"]) + 1e-9
    assert abs(float(outs["cycle"]) - float(g["metric/train_cycle_loss"])) <= 1e-5
    if "recon" in g.files:
        assert np.abs(outs["recon"].numpy() - g["recon"]).max() < 1e-5
        assert np.abs(outs["rotated_recon"].numpy() - g["rotated_recon"]).max() < 1e-5
    else:
        assert np.abs(outs["recon"][0, 0, ::8, ::8].numpy() - g["recon_b0"]).max() < 1e-5
        assert np.abs(outs["rotated_recon"][0, 0, ::8, ::8].numpy() - g["rotated_recon_b0"]).max() < 1e-5
    check_grads_against_golden(grads, g, rtol=2e-4)


def test_rvae_step_p32_matches_reference():
    _check_rvae("p32")


def test_rvae_step_p128_matches_reference():
    _check_rvae("p128")


def test_vae_step_matches_reference():
    g = load_golden("vae_step_p64.npz")
    P, L, B, seed = int(g["P"]), int(g["L"]), int(g["B"]), int(g["seed"])
    params = O.make_params(O.vae_param_shapes(P, L), seed=seed)
    x, _, _ = O.make_lattice_batch(B, P, seed=seed + 1)
    eps = torch.from_numpy(np.random.default_rng(seed + 2).standard_normal((B, L))).float()
    outs, grads = O.vae_full_step(params, x, eps, beta=1.0)
    assert abs(float(outs["loss"]) - float(g["loss"])) <= 1e-5 * float(g["loss"])
    assert np.abs(outs["recon"].numpy() - g["recon"]).max() < 1e-5
    assert rel_l2(outs["mu"], g["mu"]) < 1e-4
    check_grads_against_golden(grads, g, rtol=2e-4)


def test_stn_pretrain_matches_reference():
    g = load_golden("stn_pretrain_p32.npz")
    P, B, seed = int(g["P"]), int(g["B"]), int(g["seed"])
    params = O.make_params(O.rvae_param_shapes(P, 2), seed=seed, stn_head_std=0.5)
    x, xr, ang = O.make_lattice_batch(B, P, seed=seed + 1)
    outs, grads = O.stn_pretrain_step(params, x, xr, ang)
    assert abs(float(outs["loss"]) - float(g["loss"])) < 1e-5
    assert rel_l2(outs["theta"], g["theta"]) < 1e-5
    check_grads_against_golden(grads, g, rtol=2e-4)


def test_patch_gather_bit_exact():
    g = load_golden("patch_gather.npz")
    for HW, P in ((2048, 128), (1024, 64)):
        img = synth_image(HW, 100 + HW)
        sites = g[f"sites{HW}"]
        s3 = np.concatenate([np.zeros((len(sites), 1), dtype=np.int64), sites], 1)
        got = OP.gather_integer([img], s3, P)
        assert hashlib.sha256(got.tobytes()).hexdigest() == str(g[f"sha{HW}"])


def test_patch_gather_subpixel():
    g = load_golden("patch_gather.npz")
    img = synth_image(512, 300)
    for (cy, cx), want in zip(g["sub_sites"], g["sub_out"]):
        got = OP.gather_subpixel(img, cy, cx, 64, 8)
        assert np.abs(got - want).max() < 5e-5


def test_global_index_walk():
    assert OP.global_index_to_site([3, 0, 2], 0) == (0, 0)
    assert OP.global_index_to_site([3, 0, 2], 3) == (2, 0)
    assert OP.global_index_to_site([3, 0, 2], 4) == (2, 1)
    import pytest
    with pytest.raises(IndexError):
        OP.global_index_to_site([3, 0, 2], 5)


def replay_augment_cases():
    """(name, oracle output, golden output) for every case of tests/golden/augment.npz; the random draws
    are replayed from the recorded seeds in the reference's draw order."""
    import random

    from oracle import augment as OA
    g = load_golden("augment.npz")
    patch = synth_image(48, 400).astype(np.float32)
    for k, rot in enumerate((False, True, False, True)):
        random.seed(900 + k)
        p = OA.draw_params(rotation=rot)
        yield f"dt{k}", OA.default_transform(patch, p)[None], g[f"dt{k}"]
    for tag, P, pad in (("a", 64, 8), ("b", 32, 16)):
        img = synth_image(256, 410 + P)
        sites = g[f"sites_{tag}"]
        random.seed(1000 + P)
        for i, (cy, cx) in enumerate(sites):
            p = OA.draw_params(rotation=False)
            ang = random.uniform(0, 360)
            x, r, a = OA.paired_item(img, cy, cx, P, pad, p, ang)
            assert abs(a - g[f"pair_{tag}_angle"][i]) < 1e-12
            yield f"pair_{tag}_x{i}", x, g[f"pair_{tag}_x"][i]
            yield f"pair_{tag}_r{i}", r, g[f"pair_{tag}_r"][i]
        random.seed(2000 + P)
        for i, (cy, cx) in enumerate(sites):
            ang = random.uniform(0, 360)
            x, r, a = OA.paired_item(img, cy, cx, P, pad, None, ang)
            yield f"pairnt_{tag}_x{i}", x, g[f"pairnt_{tag}_x"][i]
            yield f"pairnt_{tag}_r{i}", r, g[f"pairnt_{tag}_r"][i]
        random.seed(3000 + P)
        for i, (cy, cx) in enumerate(sites):
            p = OA.draw_params(rotation=False)
            yield f"adapt_{tag}{i}", OA.adaptive_item(img, cy, cx, P, pad, p), g[f"adapt_{tag}"][i]


def test_augment_and_paired_rotation_match_reference():
    """a3 / paired items: float64 restatement vs the reference's torchvision fp32 grid maths.  Tolerance
    1e-4 abs on [0,1] data: two chained bilinear resamplings of white-noise images (gradient ~1/px) with
    fp32 grid rounding ~1e-5 px, amplified by the per-patch min-max."""
    n = 0
    for name, got, want in replay_augment_cases():
        assert got.shape == want.shape, name
        assert np.abs(got - want).max() < 1e-4, (name, np.abs(got - want).max())
        n += 1
    assert n == 4 + 2 * (10 + 10 + 5)


# ---- loss / batch-format branches and further shapes (tests/golden/make_golden_r3.py -> branches.npz) ---------------
class _Sub:
    """the entries of one case of branches.npz under their un-prefixed names"""

    def __init__(self, gold, tag):
        self.gold, self.pre = gold, tag + "/"
        self.files = [f[len(self.pre):] for f in gold.files if f.startswith(self.pre)]

    def __getitem__(self, k):
        return self.gold[self.pre + k]


def _branch_cases():
    from tests.golden.make_golden_r3 import RVAE_CASES, VAE_CASES
    return sorted(RVAE_CASES), sorted(VAE_CASES)


@pytest.mark.parametrize("tag", _branch_cases()[0])
def test_rvae_step_branches_match_reference(tag):
    """gamma = 0, the diversity loss, canonical_weight = 0, an unpaired batch and a pair without its angle, at latent
    sizes 2 - 16 and P = 32 / 64: the reference's own trainer (train.py:315-397, loss.py:138-186) against
    oracle.rvae.rvae_full_step"""
    from tests.golden.make_golden_r3 import RVAE_CASES, rvae_inputs
    P, L, B, seed, beta, gamma, div, cw, fmt = RVAE_CASES[tag]
    g = _Sub(load_golden("branches.npz"), tag)
    params, x, xr, ang, eps = rvae_inputs(P, L, B, seed)
    if fmt == "single":
        xr = ang = None
    elif fmt == "pair2":
        ang = None
    outs, grads = O.rvae_full_step(params, x, xr, ang, eps, beta=beta, gamma=gamma, canonical_weight=cw,
                                   use_diversity=div)
    want = float(g["metric/train_loss"])
    assert abs(float(outs["loss"]) - want) <= 1e-5 * abs(want)
    assert abs(float(outs["recon_loss"]) - float(g["metric/train_recon_loss"])) <= 1e-5 * float(g["metric/train_recon_loss"])
    assert abs(float(outs["kld"]) - float(g["metric/train_kld_loss"])) <= 1e-4 * float(g["metric/train_kld_loss"]) + 1e-9
    assert abs(float(outs["cycle"]) - float(g["metric/train_cycle_loss"])) <= 1e-5
    if gamma == 0 or fmt != "paired":
        assert float(outs["cycle"]) == 0.0 and float(g["metric/train_cycle_loss"]) == 0.0
    # 2e-4 as for the main vectors, or 3x the oracle's own fp32-vs-fp64 distance where that is larger: with two patches of
    # P = 64 the STN's gradient is a cancelling sum that fp32 itself only defines to 4.5e-4 (the reference sits 7.9e-4 away)
    floor = fp32_noise_floor(O.rvae_full_step, params, x, xr, ang, eps, beta=beta, gamma=gamma, canonical_weight=cw,
                             use_diversity=div)
    check_grads_against_golden(grads, g, rtol={k: max(2e-4, 3.0 * v) for k, v in floor.items()})


@pytest.mark.parametrize("tag", _branch_cases()[1])
def test_vae_step_shapes_match_reference(tag):
    from tests.golden.make_golden_r3 import VAE_CASES, vae_inputs
    P, L, B, seed, beta = VAE_CASES[tag]
    g = _Sub(load_golden("branches.npz"), tag)
    params, x, eps = vae_inputs(P, L, B, seed)
    outs, grads = O.vae_full_step(params, x, eps, beta=beta)
    assert abs(float(outs["loss"]) - float(g["loss"])) <= 1e-5 * float(g["loss"])
    assert abs(float(outs["recon_loss"]) - float(g["recon_loss"])) <= 1e-5 * float(g["recon_loss"])
    assert abs(float(outs["kld"]) - float(g["kld"])) <= 1e-4 * abs(float(g["kld"])) + 1e-9
    assert abs(float(outs["recon"].double().sum()) - float(g["recon_sum"])) <= 1e-5 * float(g["recon_sum"])
    assert np.abs(outs["mu"].numpy() - g["mu"]).max() < 1e-5
    check_grads_against_golden(grads, g, rtol=2e-4)


def test_latent_metric_dictionary_matches_reference():
    """livae.metrics.compute_latent_metrics is plain tensor arithmetic (no kernels): the reference's values for the same
    inputs, key for key (metrics.py:153-194).  The pixel metrics of compute_reconstruction_metrics go through the
    PSNR / SSIM kernels (checked on the GPU in tests/test_gpu_trainer_api.py and against the reference trainer's logged
    PSNR / SSIM in tests/test_gpu_dropin.py); the reference's values for these inputs are stored next to the latent ones."""
    from livae.metrics import compute_latent_metrics
    from tests.golden.make_golden_r3 import metric_inputs
    g = _Sub(load_golden("branches.npz"), "metrics")
    mu, logvar, _, _ = metric_inputs()
    got = compute_latent_metrics(mu, logvar)
    assert sorted(got) == sorted(k for k in g.files if k.startswith("latent_"))
    for k, v in got.items():
        assert abs(v - float(g[k])) <= 1e-6 * max(1.0, abs(float(g[k]))), k
