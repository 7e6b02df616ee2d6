"""micro-benchmark of individual C-ABI kernels at the C3 shapes (B=2048, P=128); used for ncu captures"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "li-vae_b200")]
import torch
from livae import ops

B = int(os.environ.get("MB", "2048"))
which = sys.argv[1:] or ["wgrad_stn2"]
dev = "cuda"
bf = torch.bfloat16


def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for w in which:
    if w == "wgrad_stn2":
        x = torch.randn(B, 64, 64, 16, device=dev).to(bf); g = torch.randn(B, 64, 64, 32, device=dev).to(bf)
        print(w, timeit(lambda: ops.tc_conv_wgrad(x, g, 5, 5, 1, 2)), "ms")
    elif w == "wgrad_d3":
        x = torch.randn(B, 66, 66, 64, device=dev).to(bf); g = torch.randn(B, 64, 64, 32, device=dev).to(bf)
        print(w, timeit(lambda: ops.tc_conv_wgrad(x, g, 3, 3, 1, 0)), "ms")
    elif w == "wgrad_d1":
        x = torch.randn(B, 18, 18, 256, device=dev).to(bf); g = torch.randn(B, 16, 16, 128, device=dev).to(bf)
        print(w, timeit(lambda: ops.tc_conv_wgrad(x, g, 3, 3, 1, 0)), "ms")
    elif w.startswith("wg_") or w.startswith("fw_") or w.startswith("dg_"):
        # generic: wg_/fw_/dg_<Cin>_<Cout>_<k>_<stride>_<pad>_<Hin>
        _, ci, co, k, st, pd, hin = w.split("_"); ci, co, k, st, pd, hin = map(int, (ci, co, k, st, pd, hin))
        ho = (hin + 2 * pd - k) // st + 1
        x = torch.randn(B, hin, hin, ci, device=dev).to(bf); g = torch.randn(B, ho, ho, co, device=dev).to(bf)
        wt = torch.randn(co, ci, k, k, device=dev)
        fl = 2.0 * B * ho * ho * ci * co * k * k
        if w.startswith("wg_"):
            t = timeit(lambda: ops.tc_conv_wgrad(x, g, k, k, st, pd))
        elif w.startswith("fw_"):
            wp = ops.tc_pack_weights(wt, co, ci, k, k, 0)
            t = timeit(lambda: ops.tc_conv(x, wp, None, k, k, st, pd, 1))
        else:
            wp = ops.tc_pack_weights(wt, co, ci, k, k, 2)
            t = timeit(lambda: ops.tc_conv_dgrad(g, wp, None, hin, hin, k, k, st, pd))
        print(f"{w}: {t:.3f} ms  {fl / t / 1e9:.0f} TF/s")
    elif w.startswith("s2blk_"):
        _, ci, co, hin = w.split("_"); ci, co, hin = int(ci), int(co), int(hin)
        g = torch.randn(B, hin // 2, hin // 2, co, device=dev).to(bf); wt = torch.randn(co, ci, 4, 4, device=dev)
        mask = torch.randn(B, hin, hin, ci, device=dev).to(bf)
        fl = 2.0 * B * (hin // 2) ** 2 * ci * co * 16
        t = timeit(lambda: ops.dgrad_s2blk(g, wt, hin, hin, relu_mask=mask))
        t2 = timeit(lambda: ops.tc_conv_dgrad(g, ops.tc_pack_weights(wt, co, ci, 4, 4, 2), None, hin, hin, 4, 4, 2, 1, relu_mask=mask))
        print(f"{w}: block form {t:.3f} ms ({fl / t / 1e9:.0f} TF/s useful)   four phases {t2:.3f} ms ({fl / t2 / 1e9:.0f} TF/s)")
    elif w == "fwd_d3":
        x = torch.randn(B, 66, 66, 64, device=dev).to(bf); wt = torch.randn(32, 64, 3, 3, device=dev)
        wp = ops.tc_pack_weights(wt, 32, 64, 3, 3, 0)
        print(w, timeit(lambda: ops.tc_conv(x, wp, None, 3, 3, 1, 0, 1)), "ms")
    elif w == "fwd_stn2":
        x = torch.randn(B, 64, 64, 16, device=dev).to(bf); wt = torch.randn(32, 16, 5, 5, device=dev)
        wp = ops.tc_pack_weights(wt, 32, 16, 5, 5, 0)
        print(w, timeit(lambda: ops.tc_conv(x, wp, None, 5, 5, 1, 2, 1)), "ms")
    elif w == "fwd_c4":
        x = torch.randn(B, 16, 16, 128, device=dev).to(bf); wt = torch.randn(256, 128, 4, 4, device=dev)
        wp = ops.tc_pack_weights(wt, 256, 128, 4, 4, 0)
        print(w, timeit(lambda: ops.tc_conv(x, wp, None, 4, 4, 2, 1, 1)), "ms")
    elif w == "thin_wgrad0":
        img = torch.rand(B, 1, 128, 128, device=dev); g = torch.randn(B, 64, 64, 16, device=dev).to(bf)
        idx = torch.randint(0, 4, (B, 64, 64, 16), device=dev, dtype=torch.uint8)
        gw = torch.empty(16, 1, 5, 5, device=dev); gb = torch.empty(16, device=dev)
        from livae._lib import call
        print(w, timeit(lambda: call("livae_thin_conv1c_wgrad", 0, img, g, idx, B, 128, 128, gw, gb)), "ms")
    elif w == "thin_wgrad_d4":
        u = torch.randn(B, 130, 130, 32, device=dev).to(bf); g = torch.randn(B, 128, 128, device=dev)
        gw = torch.empty(1, 32, 3, 3, device=dev); gb = torch.empty(1, device=dev)
        from livae._lib import call
        print(w, timeit(lambda: call("livae_thin_convc1_wgrad", u, g, B, 130, 130, gw, gb)), "ms")
    elif w == "upsample":
        from livae._lib import call
        for hw, c in ((8, 256), (16, 128), (32, 64), (64, 32)):
            x = torch.randn(B, hw, hw, c, device=dev).to(bf); u = torch.empty(B, 2 * hw + 2, 2 * hw + 2, c, device=dev, dtype=bf)
            gx = torch.empty_like(x)
            nb = (x.numel() + u.numel()) * 2
            tf = timeit(lambda: call("livae_upsample_pad_fwd_bf16", x, B, hw, hw, c, u))
            tb = timeit(lambda: call("livae_upsample_pad_bwd_bf16", u, B, hw, hw, c, x, gx))
            print(f"upsample {hw}x{hw}x{c}: fwd {tf:.3f} ms ({nb / tf / 1e6:.0f} GB/s)  bwd {tb:.3f} ms ({(nb + x.numel() * 2) / tb / 1e6:.0f} GB/s)")
    elif w == "rot":
        from livae._lib import call
        x = torch.rand(B, 1, 128, 128, device=dev); g = torch.randn(B, 1, 128, 128, device=dev)
        ang = torch.rand(B, device=dev) * 6.28
        cs = torch.stack([ang.cos(), ang.sin()], 1).contiguous()
        out = torch.empty_like(x); gi = torch.empty_like(x); gcs = torch.empty(B, 2, device=dev)
        nb = B * 128 * 128 * 4
        t = timeit(lambda: call("livae_rot_sample_fwd", x, cs, 1.0, B, 1, 128, 128, out)); print(f"rot fwd {t:.3f} ms {2 * nb / t / 1e6:.0f} GB/s")
        t = timeit(lambda: call("livae_rot_sample_bwd", x, cs, 1.0, g, B, 1, 128, 128, gi, gcs)); print(f"rot bwd full {t:.3f} ms {3 * nb / t / 1e6:.0f} GB/s")
        t = timeit(lambda: call("livae_rot_sample_bwd", x, cs, 1.0, g, B, 1, 128, 128, None, gcs)); print(f"rot bwd grid-only {t:.3f} ms {2 * nb / t / 1e6:.0f} GB/s")
