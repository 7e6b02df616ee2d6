"""time livae_upconv_c1_bwd_data alone at the C3 shape (B=2048, H=W=64); for ncu captures of that kernel"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "li-vae_b200"))
from livae._lib import call  # noqa: E402

B, H = int(os.environ.get("B", 2048)), 64
g = torch.randn(B, 2 * H, 2 * H, device="cuda")
w = torch.randn(1, 32, 3, 3, device="cuda") * 0.1
y = torch.randn(B, H, H, 32, device="cuda").clamp_min(0).to(torch.bfloat16)
gy = torch.empty_like(y)
gb = torch.empty(32, device="cuda")
for _ in range(2):
    call("livae_upconv_c1_bwd_data", g, w, y, B, H, H, gy, gb)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    call("livae_upconv_c1_bwd_data", g, w, y, B, H, H, gy, gb)
e1.record()
torch.cuda.synchronize()
print("upconv_c1_bwd_data ms:", e0.elapsed_time(e1) / 5)
