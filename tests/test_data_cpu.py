"""CPU: host-side logic of livae/data.py (no kernels are called): the random draws replay the reference's order
(checked against oracle/augment.draw_params, which is pinned to the reference's own outputs in
tests/test_oracle_golden.py), index walk / IndexError behaviour (data.py:212-220), loader batching, and the refusal to
run without a CUDA device."""
import random

import numpy as np
import pytest
import torch

from oracle import augment as OA
from oracle import patch as OP


def test_draws_replay_reference_order():
    from livae import data
    for rotation in (False, True):
        random.seed(123)
        want = [OA.draw_params(rotation=rotation) for _ in range(7)]
        random.seed(123)
        got = data.draw_transform_params(7, rotation=rotation)
        for i, w in enumerate(want):
            assert abs(float(got["scale"][i]) - w["scale"]) < 1e-7
            assert bool(got["flags"][i] & 1) == w["hflip"] and bool(got["flags"][i] & 2) == w["vflip"]
            assert tuple(got["shift"][i]) == (w["shift_y"], w["shift_x"])
            if rotation:
                assert got["angle"][i] == w["angle"]
            else:
                assert np.isnan(got["angle"][i])
    # paired items: transform draws, then the pair angle, per item
    random.seed(5)
    want = []
    for _ in range(4):
        p = OA.draw_params(rotation=False)
        want.append((p, random.uniform(0, 360)))
    random.seed(5)
    p, ang = data._draw(4, 0.5, 4, False, True, True)
    for i, (w, a) in enumerate(want):
        assert ang[i] == a and abs(float(p["scale"][i]) - w["scale"]) < 1e-7
        assert tuple(p["shift"][i]) == (w["shift_y"], w["shift_x"])
    # transform=None: only the angle is drawn
    random.seed(9)
    a0 = [random.uniform(0, 360) for _ in range(3)]
    random.seed(9)
    p, ang = data._draw(3, 0.5, 4, False, False, True)
    assert p is None and list(ang) == a0
    # jitter_amount = 0 draws no shifts (data.py:111)
    random.seed(2)
    s0 = OA.draw_params(jitter_amount=0)
    nxt = random.random()
    random.seed(2)
    g0 = data.draw_transform_params(1, jitter_amount=0)
    assert tuple(g0["shift"][0]) == (0, 0) and abs(float(g0["scale"][0]) - s0["scale"]) < 1e-7
    assert random.random() == nxt


def _same_draws(a, b):
    if a is None or b is None:
        return a is None and b is None
    if isinstance(a, dict):
        return a.keys() == b.keys() and all(_same_draws(a[k], b[k]) for k in a)
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("cfg", [(0.5, 4, False, True, True), (0.5, 4, True, True, False), (0.5, 4, False, False, True),
                                 (0.3, 0, False, True, True), (0.5, 1, True, True, True), (0.9, 100, True, True, True)])
def test_batch_draws_from_the_word_stream_equal_per_call_draws(cfg):
    """large batches parse Mersenne-Twister output words instead of calling `random` seven times per patch: the same
    numbers bit for bit (the oracle-checked per-call loop above is the yardstick) and the same generator state
    afterwards, so the reference's stream continues unchanged"""
    from livae import data
    fp, jitter, rotation, transform, pair = cfg
    for seed, n in ((0, 64), (1, 257), (2, 2048)):
        random.seed(seed); random.gauss(0, 1); random.random()            # a pending gauss value must survive
        want = data._draw_calls(n, fp, jitter, rotation, transform, pair)
        state, nxt = random.getstate(), random.random()
        random.seed(seed); random.gauss(0, 1); random.random()
        got = data._draw(n, fp, jitter, rotation, transform, pair)        # n >= 64: the stream parser
        assert _same_draws(want[0], got[0]) and _same_draws(want[1], got[1])
        assert random.getstate() == state and random.random() == nxt


def test_word_stream_draws_redraw_a_short_block(monkeypatch):
    from livae import data
    monkeypatch.setattr(data, "_DRAW_SLACK", -700)                       # first block too short by construction
    random.seed(11)
    want = data._draw_calls(128, 0.5, 4, False, True, True)
    state = random.getstate()
    random.seed(11)
    got = data._draw_stream(128, 0.5, 4, False, True, True)
    assert _same_draws(want[0], got[0]) and _same_draws(want[1], got[1]) and random.getstate() == state
    # small batches and DataLoader-worker items (n = 1) keep calling `random`
    calls = []
    monkeypatch.setattr(data, "_draw_stream", lambda *a: calls.append(a))
    data._draw(8, 0.5, 4, False, True, True)
    assert not calls


def test_source_refuses_cpu_and_foreign_transforms():
    from livae import data
    img = np.zeros((64, 64))
    with pytest.raises(RuntimeError):
        data.DevicePatchSource([img], [np.zeros((1, 2))], 32, 8, device="cpu")
    with pytest.raises(ValueError):
        data.DevicePatchSource([img], [np.zeros((1, 2))], 32, 8, transform=lambda p: p, device="cpu")
    with pytest.raises(RuntimeError):
        data.default_transform(torch.zeros(1, 40, 40))


class _FakeSource:
    """stands in for DevicePatchSource in the loader / index tests (no device needed)"""

    def __init__(self, counts):
        from livae.data import DevicePatchSource
        self._offsets = np.concatenate([[0], np.cumsum(counts)])
        self._img = np.repeat(np.arange(len(counts), dtype=np.int32), counts)
        self._yx = np.arange(2 * int(self._offsets[-1]), dtype=np.float64).reshape(-1, 2)
        self._lookup = DevicePatchSource._lookup.__get__(self)
        self.calls = []

    def __len__(self):
        return int(self._offsets[-1])

    def paired_batch(self, idx):
        self.calls.append(np.asarray(idx).copy())
        return idx


def test_index_walk_matches_reference_and_raises():
    src = _FakeSource([3, 0, 2])
    img, yx = src._lookup([0, 2, 3, 4])
    for k, i in enumerate([0, 2, 3, 4]):
        want_img, want_local = OP.global_index_to_site([3, 0, 2], i)
        assert img[k] == want_img and yx[k, 0] == 2.0 * i     # coords were laid out in global order
    with pytest.raises(IndexError):
        src._lookup([5])
    with pytest.raises(IndexError):
        src._lookup([-1])


def test_loader_batching():
    from livae.data import DevicePatchLoader
    src = _FakeSource([10, 7])
    ld = DevicePatchLoader(src, 4, mode="paired", shuffle=False, drop_last=True)
    assert len(ld) == 4
    batches = list(ld)
    assert [list(b) for b in batches] == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10, 11], [12, 13, 14, 15]]
    ld = DevicePatchLoader(src, 4, mode="paired", shuffle=False, drop_last=False)
    assert len(ld) == 5 and list(list(ld)[-1]) == [16]
    ld = DevicePatchLoader(src, 5, mode="paired", shuffle=True, drop_last=True, seed=3)
    seen = np.concatenate(list(ld))
    assert len(seen) == 15 and len(set(seen.tolist())) == 15 and not np.array_equal(seen, np.arange(15))
    ld2 = DevicePatchLoader(src, 5, mode="paired", shuffle=True, drop_last=True, seed=3)
    assert np.array_equal(np.concatenate(list(ld2)), seen)           # seeded
    with pytest.raises(ValueError):
        DevicePatchLoader(src, 4, mode="nope")
