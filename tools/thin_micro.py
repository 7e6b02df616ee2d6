"""Times the eight thin-layer kernels at the C3 shapes (B=2048, P=128) in both modes (tcgen05 / SIMT) and
prints the relative L2 difference between the two implementations' outputs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "li-vae_b200")]
import torch
from livae import _lib
from livae._lib import call

B = int(os.environ.get("MB", "2048"))
P = int(os.environ.get("MP", "128"))
dev, bf = "cuda", torch.bfloat16
L = _lib.lib()
only = sys.argv[1:]


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-30))


torch.manual_seed(0)
h = P // 2
img = torch.rand(B, 1, P, P, device=dev)
cases = {}

# --- STN conv1 fwd / wgrad
w0 = torch.randn(16, 1, 5, 5, device=dev) * 0.2; b0 = torch.randn(16, device=dev) * 0.1
a1 = torch.empty(B, h, h, 16, dtype=bf, device=dev); idx1 = torch.empty(B, h, h, 16, dtype=torch.uint8, device=dev)
cases["conv1_fwd"] = (lambda: call("livae_thin_conv1c_fwd", 0, img, w0, b0, B, P, P, a1, idx1), lambda: (a1.clone(),),
                      B * (P * P * 4 + h * h * 16 * 3))
gp = (torch.randn(B, h, h, 16, device=dev)).to(bf)
idxr = torch.randint(0, 4, (B, h, h, 16), device=dev, dtype=torch.uint8)
gw0 = torch.empty(16, 1, 5, 5, device=dev); gb0 = torch.empty(16, device=dev)
cases["conv1_wgrad"] = (lambda: call("livae_thin_conv1c_wgrad", 0, img, gp, idxr, B, P, P, gw0, gb0),
                        lambda: (gw0.clone(), gb0.clone()), B * (P * P * 4 + h * h * 16 * 3))
# --- encoder c1 fwd / wgrad / dgrad
w1 = torch.randn(32, 1, 4, 4, device=dev) * 0.25; b1 = torch.randn(32, device=dev) * 0.1
h1 = torch.empty(B, h, h, 32, dtype=bf, device=dev)
cases["c1_fwd"] = (lambda: call("livae_thin_conv1c_fwd", 1, img, w1, b1, B, P, P, h1, None), lambda: (h1.clone(),),
                   B * (P * P * 4 + h * h * 32 * 2))
g1 = torch.randn(B, h, h, 32, device=dev).to(bf)
gw1 = torch.empty(32, 1, 4, 4, device=dev); gb1 = torch.empty(32, device=dev)
cases["c1_wgrad"] = (lambda: call("livae_thin_conv1c_wgrad", 1, img, g1, None, B, P, P, gw1, gb1),
                     lambda: (gw1.clone(), gb1.clone()), B * (P * P * 4 + h * h * 32 * 2))
gx1 = torch.empty(B, 1, P, P, device=dev)
cases["c1_dgrad"] = (lambda: call("livae_thin_conv1c_dgrad", g1, w1, B, P, P, gx1), lambda: (gx1.clone(),),
                     B * (P * P * 4 + h * h * 32 * 2))
# --- decoder d4 fwd / wgrad / dgrad
u4 = torch.randn(B, P + 2, P + 2, 32, device=dev).to(bf)
w4 = torch.randn(1, 32, 3, 3, device=dev) * 0.1; b4 = torch.tensor([0.05], device=dev)
rec = torch.empty(B, 1, P, P, device=dev)
cases["d4_fwd"] = (lambda: call("livae_thin_convc1_fwd", u4, w4, b4, B, P + 2, P + 2, 2, rec), lambda: (rec.clone(),),
                   B * ((P + 2) ** 2 * 64 + P * P * 4))
gpre = torch.randn(B, 1, P, P, device=dev)
gw4 = torch.empty(1, 32, 3, 3, device=dev); gb4 = torch.empty(1, device=dev)
cases["d4_wgrad"] = (lambda: call("livae_thin_convc1_wgrad", u4, gpre, B, P + 2, P + 2, gw4, gb4),
                     lambda: (gw4.clone(), gb4.clone()), B * ((P + 2) ** 2 * 64 + P * P * 4))
gu = torch.empty(B, P + 2, P + 2, 32, dtype=bf, device=dev)
cases["d4_dgrad"] = (lambda: call("livae_thin_conv1c_fwd", 2, gpre, w4, None, B, P, P, gu, None), lambda: (gu.clone(),),
                     B * ((P + 2) ** 2 * 64 + P * P * 4))

for name, (fn, get, nbytes) in cases.items():
    if only and name not in only:
        continue
    L.livae_thin_set_tc(0)
    t_simt = timeit(fn); ref = get()
    L.livae_thin_set_tc(1)
    t_tc = timeit(fn); got = get()
    errs = " ".join(f"{rel(g, r):.2e}" for g, r in zip(got, ref))
    print(f"{name:12s} simt {t_simt:7.3f} ms   tc {t_tc:7.3f} ms  ({nbytes / t_tc / 1e6:7.0f} GB/s)   rel diff {errs}", flush=True)
