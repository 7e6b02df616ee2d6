"""CPU: host-side logic of livae/train.py that involves no kernel: batch unpacking (reference train.py:316-339), the
post-clip gradient norm the metric block reports (train.py:396-405), the device-side metric accumulator's averaging, and
the prefetcher's pass-through on a non-CUDA device."""
import math

import pytest
import torch


def test_unpack_rvae_batch_formats():
    from livae.train import _unpack_rvae_batch
    cpu = torch.device("cpu")
    x, xr = torch.rand(3, 1, 8, 8), torch.rand(3, 1, 8, 8)
    ang = torch.tensor([0.1, 0.2, 0.3], dtype=torch.float64)
    a, b, c = _unpack_rvae_batch((x, xr, ang), cpu)                   # paired dataset, default collate
    assert a is x and b is xr and c.dtype == torch.float32 and torch.allclose(c, ang.float())
    a, b, c = _unpack_rvae_batch([x, xr, [0.5, 1.5, 2.5]], cpu)       # python-float angles
    assert c.dtype == torch.float32 and c.tolist() == [0.5, 1.5, 2.5]
    a, b, c = _unpack_rvae_batch((x, xr), cpu)                        # pair without angle
    assert a is x and b is xr and c is None
    a, b, c = _unpack_rvae_batch((x,), cpu)                           # TensorDataset item
    assert a is x and b is None and c is None
    a, b, c = _unpack_rvae_batch(x, cpu)                              # bare tensor
    assert a is x and b is None and c is None

    class Recipe:                                                     # a RecipeBatch that reached the loop un-pinned
        def materialise(self):
            return [x, xr, ang]
    a, b, c = _unpack_rvae_batch(Recipe(), cpu)
    assert a is x and b is xr and c.dtype == torch.float32


@pytest.mark.parametrize("scale", [0.01, 1.0, 300.0])
def test_post_clip_norm_is_what_the_reference_measures(scale):
    """the reference clips, then measures the norm of what is left (train.py:396-405); the step hands back the pre-clip
    norm (what clip_grad_norm_ returns) and the metric block derives the same number from it"""
    from livae.train import _post_clip_norm
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(5, 7)), torch.nn.Parameter(torch.randn(11))]
    for p in ps:
        p.grad = torch.randn_like(p) * scale
    max_norm = 20.0
    pre = torch.nn.utils.clip_grad_norm_(ps, max_norm=max_norm)
    measured = math.sqrt(sum(float(p.grad.norm()) ** 2 for p in ps))   # train.py:399-405
    got = float(_post_clip_norm(pre, max_norm))
    assert got == pytest.approx(measured, rel=1e-5)
    assert got <= max_norm * (1 + 1e-6)


def test_metric_accumulator_averages_and_copies_its_first_value():
    from livae.train import _DevAccum
    acc = _DevAccum(torch.device("cpu"))
    static = torch.tensor(2.0)                  # stands for a view of static CUDA-graph memory
    acc.add(loss=static, n=1)
    static.fill_(100.0)                         # the next replay overwrites it: the accumulated value must not follow
    acc.add(loss=torch.tensor(4.0), n=3.0)
    assert acc.averages(2) == {"loss": 3.0, "n": 2.0}
    assert _DevAccum(torch.device("cpu")).averages(1) == {}


def test_prefetcher_passes_through_off_cuda():
    from livae.train import DevicePrefetcher
    batches = [(torch.full((2, 2), float(i)), [0.1 * i, 0.2 * i]) for i in range(4)]
    pf = DevicePrefetcher(batches, torch.device("cpu"))
    assert len(pf) == 4
    out = list(pf)
    assert all(o[0] is b[0] and o[1] == b[1] for o, b in zip(out, batches))
