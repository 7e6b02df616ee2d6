import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "li-vae_b200")]
import numpy as np, torch, torch.nn.functional as F
from livae import ops, _lib
L = _lib.lib()
def bf(t): return t.to(torch.bfloat16).float()
def nhwc(t): return t.permute(0, 2, 3, 1).contiguous()
cases = [(2, 64, 18, 18, 32, 3, 1, 0), (2, 32, 20, 20, 64, 3, 1, 1), (2, 16, 16, 16, 32, 5, 1, 2), (2, 128, 18, 18, 64, 3, 1, 0),
         (2, 32, 32, 32, 64, 4, 2, 1), (2, 64, 32, 32, 128, 4, 2, 1)]
for mode in (0, 1, 2, 3):
    L.livae_tc_set_halo_mode(mode)
    res = []
    for (B, Ci, H, W, Co, k, s, p) in cases:
        rng = np.random.default_rng(1)
        x = bf(torch.tensor(rng.standard_normal((B, Ci, H, W)).astype(np.float32)))
        w = bf(torch.tensor((rng.standard_normal((Co, Ci, k, k)) / np.sqrt(Ci * k * k)).astype(np.float32)))
        want = nhwc(F.conv2d(x, w, None, stride=s, padding=p))
        wp = ops.tc_pack_weights(w.cuda(), Co, Ci, k, k, 0)
        y = ops.tc_conv(nhwc(x).cuda().to(torch.bfloat16), wp, None, k, k, s, p, 0, out_f32=True).cpu()
        res.append(float((y - want).norm() / want.norm()))
    print("mode", mode, " ".join(f"{r:.1e}" for r in res))
