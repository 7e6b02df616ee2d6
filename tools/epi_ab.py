"""A/B of the coalescing (shared-memory staged) bf16 epilogue of the halo convolution kernel against the direct one
(livae_tc_set_halo_mode 1 vs 3) on the C3 layer shapes.  usage: python tools/epi_ab.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "li-vae_b200")]
import torch
from livae import _lib, ops

L = _lib.lib()
B, dev, bf = 2048, "cuda", torch.bfloat16


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


cases = [("fw", 64, 32, 3, 1, 0, 66, False), ("dg", 64, 32, 3, 1, 0, 66, False), ("fw", 128, 64, 3, 1, 0, 34, False),
         ("dg", 128, 64, 3, 1, 0, 34, False), ("fw", 256, 128, 3, 1, 0, 18, False), ("dg", 256, 128, 3, 1, 0, 18, False),
         ("fw", 32, 64, 4, 2, 1, 64, False), ("fw", 64, 128, 4, 2, 1, 32, False), ("fw", 128, 256, 4, 2, 1, 16, False),
         ("dg", 64, 128, 4, 2, 1, 32, True), ("dg", 128, 256, 4, 2, 1, 16, True)]
for kind, ci, co, k, st, pd, hin, masked in cases:
    ho = (hin + 2 * pd - k) // st + 1
    wt = torch.randn(co, ci, k, k, device=dev)
    x = torch.randn(B, hin, hin, ci, device=dev).to(bf)
    g = torch.randn(B, ho, ho, co, device=dev).to(bf)
    mask = torch.randn(B, hin, hin, ci, device=dev).to(bf) if masked else None
    if kind == "fw":
        wp = ops.tc_pack_weights(wt, co, ci, k, k, 0)
        fn = lambda: ops.tc_conv(x, wp, None, k, k, st, pd, 1)
    else:
        wp = ops.tc_pack_weights(wt, co, ci, k, k, 2)
        fn = lambda: ops.tc_conv_dgrad(g, wp, None, hin, hin, k, k, st, pd, relu_mask=mask)
    res = {}
    outs = {}
    for mode in (3, 1):
        L.livae_tc_set_halo_mode(mode)
        res[mode] = timeit(fn)
        outs[mode] = fn().float()
    L.livae_tc_set_halo_mode(1)
    same = bool(torch.equal(outs[1], outs[3]))
    fl = 2.0 * B * ho * ho * ci * co * k * k
    print(f"{kind} {ci}->{co} k{k} s{st} {hin}x{hin}{' mask' if masked else ''}: direct {res[3]:.3f} ms  staged {res[1]:.3f} ms "
          f"({fl / res[1] / 1e9:.0f} TF/s)  identical={same}")
