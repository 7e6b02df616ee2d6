import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "li-vae_b200")]
import torch, torch.nn.functional as F
from livae import _lib
from livae._lib import call
L = _lib.lib()
dev, bf, P = "cuda", torch.bfloat16, 128
kind = sys.argv[1]
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
for B in [int(v) for v in sys.argv[2:]]:
    torch.manual_seed(B)
    if kind == "d4":
        u4 = torch.randn(B, P + 2, P + 2, 32, device=dev).to(bf); gpre = torch.randn(B, 1, P, P, device=dev)
        ref = torch.nn.grad.conv2d_weight(u4.permute(0, 3, 1, 2).double(), (1, 32, 3, 3), gpre.double())
        out = []
        for mode in (0, 1, 1):
            L.livae_thin_set_tc(mode)
            gw = torch.empty(1, 32, 3, 3, device=dev); gb = torch.empty(1, device=dev)
            call("livae_thin_convc1_wgrad", u4, gpre, B, P + 2, P + 2, gw, gb); torch.cuda.synchronize()
            out.append(rel(gw, ref))
        print(kind, B, "simt %.2e tc %.2e tc %.2e" % tuple(out), flush=True)
    else:
        img = torch.rand(B, 1, P, P, device=dev); g = torch.randn(B, 64, 64, 32, device=dev).to(bf)
        ref = torch.nn.grad.conv2d_weight(img.double(), (32, 1, 4, 4), g.permute(0, 3, 1, 2).double(), stride=2, padding=1)
        out = []
        for mode in (0, 1, 1):
            L.livae_thin_set_tc(mode)
            gw = torch.empty(32, 1, 4, 4, device=dev); gb = torch.empty(32, device=dev)
            call("livae_thin_conv1c_wgrad", 1, img, g, None, B, P, P, gw, gb); torch.cuda.synchronize()
            out.append(rel(gw, ref))
        print(kind, B, "simt %.2e tc %.2e tc %.2e" % tuple(out), flush=True)
