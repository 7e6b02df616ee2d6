"""Pipeline timeline of one traced kernel launch (livae_set_probe): per warp role, the clock64 deltas
between its pipeline points for CTA 0.  usage: python tools/probe.py <case> [first_record]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "li-vae_b200")]
import torch
from livae import _lib, ops
L = _lib.lib()
B, dev, bf = int(os.environ.get("MB", "2048")), "cuda", torch.bfloat16
case = sys.argv[1]
first = int(sys.argv[2]) if len(sys.argv) > 2 else 100
if case == "fwd_stn2":
    x = torch.randn(B, 64, 64, 16, device=dev).to(bf); wp = ops.tc_pack_weights(torch.randn(32, 16, 5, 5, device=dev), 32, 16, 5, 5, 0)
    run = lambda: ops.tc_conv(x, wp, None, 5, 5, 1, 2, 1)
elif case == "fwd_d3":
    x = torch.randn(B, 66, 66, 64, device=dev).to(bf); wp = ops.tc_pack_weights(torch.randn(32, 64, 3, 3, device=dev), 32, 64, 3, 3, 0)
    run = lambda: ops.tc_conv(x, wp, None, 3, 3, 1, 0, 1)
elif case == "fwd_d1":
    x = torch.randn(B, 18, 18, 256, device=dev).to(bf); wp = ops.tc_pack_weights(torch.randn(128, 256, 3, 3, device=dev), 128, 256, 3, 3, 0)
    run = lambda: ops.tc_conv(x, wp, None, 3, 3, 1, 0, 1)
elif case == "fwd_c3":
    x = torch.randn(B, 32, 32, 64, device=dev).to(bf); wp = ops.tc_pack_weights(torch.randn(128, 64, 4, 4, device=dev), 128, 64, 4, 4, 0)
    run = lambda: ops.tc_conv(x, wp, None, 4, 4, 2, 1, 1)
if case == "c5p_fwd":
    x = torch.randn(B, 64, 64, 16, device=dev).to(bf); w = torch.randn(32, 16, 5, 5, device=dev); bb = torch.randn(32, device=dev)
    run = lambda: ops.conv5pool_fwd(x, w, bb)
elif case == "s2blk_c2":
    g = torch.randn(B, 32, 32, 64, device=dev).to(bf); w = torch.randn(64, 32, 4, 4, device=dev)
    run = lambda: ops.dgrad_s2blk(g, w, 64, 64)
elif case == "s2blk_c3":
    g = torch.randn(B, 16, 16, 128, device=dev).to(bf); w = torch.randn(128, 64, 4, 4, device=dev)
    run = lambda: ops.dgrad_s2blk(g, w, 32, 32)
if case == "wgrad_stn2":
    x = torch.randn(B, 64, 64, 16, device=dev).to(bf); g = torch.randn(B, 64, 64, 32, device=dev).to(bf)
    run = lambda: ops.tc_conv_wgrad(x, g, 5, 5, 1, 2)
elif case == "wgrad_d3":
    x = torch.randn(B, 66, 66, 64, device=dev).to(bf); g = torch.randn(B, 64, 64, 32, device=dev).to(bf)
    run = lambda: ops.tc_conv_wgrad(x, g, 3, 3, 1, 0)
elif case == "wgrad_d1":
    x = torch.randn(B, 18, 18, 256, device=dev).to(bf); g = torch.randn(B, 16, 16, 128, device=dev).to(bf)
    run = lambda: ops.tc_conv_wgrad(x, g, 3, 3, 1, 0)
if case.startswith("up_"):
    # phase-folded decoder blocks (csrc/upfold.cu): up_<fwd|dg|wg>_<d2|d3>
    _, what, layer = case.split("_")
    h, ci, co = {"d2": (16, 128, 64), "d3": (32, 64, 32)}[layer]
    x = torch.randn(B, h, h, ci, device=dev).clamp_min(0).to(bf); wt = torch.randn(co, ci, 3, 3, device=dev) * 0.05
    bias = torch.zeros(co, device=dev)
    gz = torch.randn(B, 2 * h, 2 * h, co, device=dev).to(bf)
    wf, wd = ops.upfold_pack(wt)
    ctb = torch.zeros(2 * B, 2, 2 * h, co, device=dev); clr = torch.zeros(2 * B, 2, 2 * h, co, device=dev)
    y = torch.empty(B, 2 * h, 2 * h, co, device=dev, dtype=bf); gx = torch.empty_like(x)
    from livae._lib import call
    if what == "fwd":
        run = lambda: call("livae_upfold_fwd", x, wf, bias, B, h, h, ci, co, y)
    elif what == "dg":
        run = lambda: call("livae_upfold_dgrad", gz, wd, x, B, h, h, ci, co, gx)
    else:
        gw = torch.empty_like(wt); ws = torch.empty(L.livae_upfold_wgrad_ws_bytes(ci, co) // 4, device=dev)
        run = lambda: call("livae_upfold_wgrad", x, gz, None, None, B, h, h, ci, co, gw, ws)
if case.startswith("dg_") or case.startswith("fw_"):
    # generic: dg_/fw_<Cin>_<Cout>_<k>_<stride>_<pad>_<Hin>
    _, ci, co, k, st, pd, hin = case.split("_"); ci, co, k, st, pd, hin = map(int, (ci, co, k, st, pd, hin))
    ho = (hin + 2 * pd - k) // st + 1
    wt = torch.randn(co, ci, k, k, device=dev)
    if case.startswith("dg_"):
        g = torch.randn(B, ho, ho, co, device=dev).to(bf); wp = ops.tc_pack_weights(wt, co, ci, k, k, 2)
        run = lambda: ops.tc_conv_dgrad(g, wp, None, hin, hin, k, k, st, pd)
    else:
        x = torch.randn(B, hin, hin, ci, device=dev).to(bf); wp = ops.tc_pack_weights(wt, co, ci, k, k, 0)
        run = lambda: ops.tc_conv(x, wp, None, k, k, st, pd, 1)
run(); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print(f"--- {case}: {e0.elapsed_time(e1):.3f} ms untraced")
buf = torch.zeros(4096, dtype=torch.int64, device=dev)
L.livae_set_probe(buf.data_ptr())
run(); torch.cuda.synchronize()
L.livae_set_probe(None)
r = buf.cpu().tolist()
spans = sorted(v for v in r[3 * 1024:4 * 1024] if v != 0)
if spans:
    print(f"--- {case} per-CTA spans ({len(spans)} CTAs): min {spans[0]}  median {spans[len(spans) // 2]}  max {spans[-1]} cycles")
for role, name in enumerate(("producer", "mma", "epilogue")):
    recs = [(v >> 56, v & ((1 << 56) - 1)) for v in r[role * 1024:(role + 1) * 1024] if v != 0]
    if not recs:
        continue
    print(f"--- {case} {name}: {len(recs)} records; showing from #{first}")
    prev = recs[first - 1][1] if first > 0 and len(recs) > first else recs[0][1]
    print(f"  total span {recs[-1][1] - recs[0][1]} cycles over {len(recs)} records")
    for i, (s, t) in enumerate(recs[first:first + 24]):
        print(f"  slot {s}  dt={t - prev:6d}")
        prev = t
