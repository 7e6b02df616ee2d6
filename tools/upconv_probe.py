"""time the fused decoder-d4 kernels alone at the C3 shape (B=2048, H=W=64); also the target of ncu captures"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "li-vae_b200"))
from livae._lib import call  # noqa: E402

B, H = int(os.environ.get("B", 2048)), 64
g = torch.randn(B, 2 * H, 2 * H, device="cuda")
w = torch.randn(1, 32, 3, 3, device="cuda") * 0.1
bias = torch.zeros(1, device="cuda")
x = torch.randn(B, H, H, 32, device="cuda").clamp_min(0).to(torch.bfloat16)
gx = torch.empty_like(x)
gbl, gw, gb = torch.empty(32, device="cuda"), torch.empty(1, 32, 3, 3, device="cuda"), torch.empty(1, device="cuda")
out = torch.empty(B, 1, 2 * H, 2 * H, device="cuda")


def bwd():
    call("livae_upconv_c1_bwd", g, w, x, B, H, H, gx, gbl, gw, gb)


def fwd():
    call("livae_upconv_c1_fwd", x, w, bias, B, H, H, 2, out)


for name, fn, nbytes in (("upconv_c1_fwd", fwd, x.numel() * 2 + out.numel() * 4),
                         ("upconv_c1_bwd", bwd, g.numel() * 4 + 2 * x.numel() * 2)):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name}: {ms:.3f} ms, {nbytes / ms / 1e6:.0f} GB/s algorithmic")
