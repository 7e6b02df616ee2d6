"""GPU: livae.optim.FlatAdamW against torch.optim.AdamW / Adam -- parameter groups (--stn-lr), parameters frozen after
construction (--freeze-stn, scripts/train_rvae.py:143-159, 184-189), and checkpoint round trips in torch's format."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _models():
    torch.manual_seed(0)
    a = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3), torch.nn.Linear(3, 2)).cuda()
    return a, copy.deepcopy(a)


def _loss(m, x):
    return (m(x) ** 2).sum()


def _groups(m, lr_a=1e-2, lr_b=3e-3):
    return [{"params": list(m[0].parameters()), "lr": lr_a},
            {"params": list(m[2].parameters()) + list(m[3].parameters()), "lr": lr_b, "weight_decay": 0.1}]


def _close(a, b, tol=2e-6):
    for p, q in zip(a.parameters(), b.parameters()):
        assert torch.allclose(p, q, rtol=tol, atol=tol), float((p - q).abs().max())


def test_groups_and_frozen_parameters_match_torch():
    from livae.optim import FlatAdamW
    ma, mb = _models()
    oa = torch.optim.AdamW(_groups(ma), lr=1e-3, weight_decay=1e-2)
    ob = FlatAdamW(_groups(mb), lr=1e-3, weight_decay=1e-2)
    assert len(ob.param_groups) == 2 and ob.param_groups[1]["weight_decay"] == 0.1
    x = torch.randn(16, 7, device="cuda")
    for step in range(6):
        if step == 2:            # freeze the middle layer AFTER the optimisers exist
            for m in (ma, mb):
                for p in m[2].parameters():
                    p.requires_grad = False
        if step == 4:
            for m in (ma, mb):
                for p in m[2].parameters():
                    p.requires_grad = True
        frozen_before = [p.detach().clone() for p in mb[2].parameters()]
        for m, o in ((ma, oa), (mb, ob)):
            o.zero_grad(set_to_none=True)
            _loss(m, x).backward()
            o.step()
        if step in (2, 3):       # no weight decay, no moment update on a parameter without gradient
            for p, q in zip(mb[2].parameters(), frozen_before):
                assert torch.equal(p, q)
        if step < 4:             # (after the release torch's per-parameter step counts lag ours by design)
            _close(ma, mb)
    assert all(p.grad is not None and p.grad.data_ptr() != 0 for p in mb.parameters())


def test_adam_coupled_and_single_group():
    from livae.optim import FlatAdamW
    ma, mb = _models()
    oa = torch.optim.Adam(ma.parameters(), lr=1e-2, weight_decay=0.05)
    ob = FlatAdamW(mb.parameters(), lr=1e-2, weight_decay=0.05, decoupled=False)
    x = torch.randn(16, 7, device="cuda")
    for _ in range(4):
        for m, o in ((ma, oa), (mb, ob)):
            o.zero_grad()
            _loss(m, x).backward()
            o.step()
    _close(ma, mb)


def test_state_dict_round_trip_and_torch_interchange():
    from livae.optim import FlatAdamW
    ma, mb = _models()
    oa = torch.optim.AdamW(ma.parameters(), lr=1e-2, weight_decay=1e-2)
    ob = FlatAdamW(mb.parameters(), lr=1e-2, weight_decay=1e-2)
    x = torch.randn(16, 7, device="cuda")
    for _ in range(3):
        for m, o in ((ma, oa), (mb, ob)):
            o.zero_grad()
            _loss(m, x).backward()
            o.step()
    sd = ob.state_dict()
    ref = oa.state_dict()
    assert set(sd["state"]) == set(ref["state"]) and float(sd["state"][0]["step"]) == 3.0
    for i in ref["state"]:
        assert torch.allclose(sd["state"][i]["exp_avg"], ref["state"][i]["exp_avg"], rtol=1e-3, atol=1e-6)
        assert torch.allclose(sd["state"][i]["exp_avg_sq"], ref["state"][i]["exp_avg_sq"], rtol=1e-3, atol=1e-8)
    # resume: a fresh FlatAdamW loaded from torch's checkpoint continues exactly like torch does
    mc = copy.deepcopy(ma)
    oc = FlatAdamW(mc.parameters(), lr=5.0, weight_decay=0.7)
    oc.load_state_dict(copy.deepcopy(ref))
    assert oc.param_groups[0]["lr"] == 1e-2 and oc.param_groups[0]["weight_decay"] == 1e-2
    for m, o in ((ma, oa), (mc, oc)):
        o.zero_grad()
        _loss(m, x).backward()
        o.step()
    _close(ma, mc)
    # and the other way: torch's AdamW accepts ours
    od = torch.optim.AdamW(copy.deepcopy(mb).parameters(), lr=1e-2, weight_decay=1e-2)
    od.load_state_dict(sd)


def test_tensor_core_engine_takes_odd_latent_dims():
    """2*latent_dim = 40 columns pad to 64, not 48 (the backward kernels take 16, 32 or multiples of 64)"""
    import livae
    livae.set_engine("tc")
    m = livae.RVAE(latent_dim=20, in_channels=1, patch_size=32).cuda()
    x = torch.rand(4, 1, 32, 32, device="cuda")
    out = m(x)
    (out[0].sum() + out[3].sum() + out[4].sum()).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    v = livae.VAE(latent_dim=20, in_channels=1, patch_size=32).cuda()
    r, mu, lv = v(x)
    (r.sum() + mu.sum() + lv.sum()).backward()
    assert mu.shape == (4, 20) and all(torch.isfinite(p.grad).all() for p in v.parameters())
