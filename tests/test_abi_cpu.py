"""CPU: the C-ABI library builds/loads and exports every symbol include/livae_b200.h declares;
the product refuses CPU tensors (no fallback).  No compute calls."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "livae_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(livae_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def libpath():
    import importlib.util
    spec = importlib.util.spec_from_file_location("livae_build", os.path.join(ROOT, "li-vae_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def test_header_symbols_exported(libpath):
    lib = ctypes.CDLL(libpath)
    names = _declared()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.livae_abi_version.restype = ctypes.c_int
    assert lib.livae_abi_version() >= 1


def test_binding_table_matches_header(libpath):
    from livae import _lib
    assert sorted(_lib.exported_symbols()) == _declared()
    _lib.lib()   # argtypes resolve for every entry


def test_binding_arity_matches_header(libpath):
    """every ctypes signature in livae/_lib.py has as many arguments, of the same pointer / scalar kind, as the
    prototype in include/livae_b200.h (a wrong table is a host-side crash, not an error code)"""
    from livae import _lib
    src = open(os.path.join(ROOT, "include", "livae_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = dict(re.findall(r"\bint\s+(livae_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src))
    bad = []
    for name, sig in _lib._SIGS.items():
        assert name in protos, name
        args = [a.strip() for a in protos[name].split(",") if a.strip()]
        kinds = ""
        for a in args:
            if "livae_stream_t" in a:
                kinds += "s"
            elif "*" in a:
                kinds += "p"
            elif re.match(r"(const\s+)?(int64_t|long long)\b", a):
                kinds += "l"
            elif re.match(r"(const\s+)?(float|double)\b", a):
                kinds += "f"
            else:
                kinds += "i"
        want = sig.replace("d", "p").replace("t", "p")
        if kinds != want:
            bad.append((name, sig, kinds))
    assert not bad, bad


def test_no_cpu_fallback(libpath):
    import livae
    from livae import ops
    with pytest.raises(RuntimeError):
        ops.rot_sample(torch.zeros(1, 1, 8, 8), torch.zeros(1, 2), 1.0)
    m = livae.RVAE(2, 1, 32)
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 1, 32, 32))
    with pytest.raises(RuntimeError):
        ops.patch_gather(torch.zeros(1, 8, 8), torch.zeros((1, 3), dtype=torch.int32), 4)


def test_state_dict_keys_match_reference_layout():
    """checkpoint compatibility"""
    import livae
    from oracle import rvae as O
    m = livae.RVAE(2, 1, 128)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == O.rvae_param_shapes(128, 2)
    v = livae.VAE(16, 1, 64)
    assert {k: tuple(t.shape) for k, t in v.state_dict().items()} == O.vae_param_shapes(64, 16)


def test_canonical_cache_is_one_shot_and_identity_keyed():
    """Encoder.take_canonical (the STN's rotated batch reused as rotate_to_canonical(x, theta), train.py:389): only
    for exactly the tensors of the last forward, handed out once, and never left on the module afterwards."""
    import copy

    import livae
    enc = livae.Encoder(1, 2, 32)
    x, th, xr = torch.zeros(2, 1, 32, 32), torch.zeros(2, 1), torch.ones(2, 1, 32, 32)
    assert enc.take_canonical(x, th) is None                      # nothing cached yet
    enc._canonical = (x, th, xr)
    assert enc.take_canonical(x.clone(), th) is None              # a different tensor object: refused, cache dropped
    assert enc._canonical is None
    enc._canonical = (x, th, xr)
    assert enc.take_canonical(x, th) is xr and enc.take_canonical(x, th) is None
    copy.deepcopy(enc)                                            # an idle module carries no batch / graph


def test_halo_tile_geometry_chooser(libpath):
    """host logic of the tcgen05 halo convolution (csrc/conv_tc.cu launch_conv_tc_halo): the tile is 128 consecutive
    positions of a bw-wide box, bw chosen in 16..20 to cover the output grid with the fewest tiles"""
    import ctypes as C

    from livae import _lib
    L = _lib.lib()

    def geo(H, W, sx):
        out = [C.c_int() for _ in range(4)]
        assert L.livae_tc_halo_geometry(H, W, sx, *[C.byref(o) for o in out]) == 0
        return tuple(o.value for o in out)

    def tiles_at(H, W, sx, bw):
        tw, th = bw - sx, 128 // bw
        return -(-W // tw) * -(-H // th)

    # the cases the round-2 notes quote: 32 x 32 and 16 x 16 maps under 3x3 taps, the 64-wide strips
    assert geo(32, 32, 2) == (18, 7, 16, 10) and tiles_at(32, 32, 2, 16) == 12
    assert geo(16, 16, 2)[3] == 3 and tiles_at(16, 16, 2, 16) == 4
    assert geo(64, 64, 2)[3] <= tiles_at(64, 64, 2, 16) == 40
    for H in (4, 8, 9, 16, 31, 32, 64, 100, 16382):
        for W in (8, 14, 16, 28, 32, 33, 64, 66):
            for sx in (0, 1, 2, 4):
                bw, th, tw, tiles = geo(H, W, sx)
                assert 16 <= bw <= 20 and th == 128 // bw and tw == bw - sx and th * bw <= 128
                assert tiles == tiles_at(H, W, sx, bw) == min(tiles_at(H, W, sx, b) for b in range(16, 21))
                assert -(-W // tw) * tw >= W and -(-H // th) * th >= H          # the tiles cover the grid
    assert L.livae_tc_halo_geometry(0, 8, 2, None, None, None, None) != 0          # argument errors are reported, not crashed on


def test_benchmarked_layers_are_eligible_for_the_tensor_core_kernels(libpath):
    """host arithmetic behind the C ABI (no device work): every GEMM-shaped layer of the benchmarked rVAE (C3: P = 128,
    B = 2048; reference model.py:185-237, 262-300, 330-372) is accepted by the tcgen05 kernels' eligibility checks, so the
    `tc` engine cannot silently run the exact SIMT path there; shapes outside the kernels' channel rules are refused;
    the weight-gradient workspace is the fp32 gradient itself"""
    import ctypes as C
    from livae import _lib
    L = _lib.lib()

    def tc(B, H, W, Ci, Co, k, s, p):
        d = _lib.TcConvDesc(B, H, W, Ci, Co, k, k, s, p, 0, 0)
        return L.livae_tc_conv_supported(C.byref(d)), L.livae_tc_wgrad_ws_bytes(C.byref(d))

    for B in (2048, 256, 8):
        # encoder c2 - c4: Conv2d(k 4, s 2, p 1); decoder d1, d2: 3 x 3 over the reflection-padded up-sampled map
        for H, Ci, Co, k, s, p in ((64, 32, 64, 4, 2, 1), (32, 64, 128, 4, 2, 1), (16, 128, 256, 4, 2, 1),
                                   (18, 256, 128, 3, 1, 0), (34, 128, 64, 3, 1, 0), (66, 64, 32, 3, 1, 0)):
            ok, ws = tc(B, H, H, Ci, Co, k, s, p)
            assert ok == 1 and ws == Co * Ci * k * k * 4, (B, H, Ci, Co)
        assert L.livae_tc_conv5pool_supported(B, 64, 64, 16, 32) == 1          # STN conv2 + pool on the 64 x 64 pooled map
        assert L.livae_upfold_supported(B, 32, 32, 64, 32) == 1                 # d3 folded onto its 32 x 32 input
        assert L.livae_tc_dgrad_s2blk_supported(64, 64, 32, 64) == 1            # encoder c2 data gradient, block form
    assert L.livae_tc_conv5pool_wgrad_ws_bytes(16, 32) >= 32 * 16 * 25 * 4
    assert L.livae_upfold_wgrad_ws_bytes(64, 32) >= 32 * 64 * 9 * 4
    # refused: channel counts the MMA tiling has no layout for, and the 5-channel case of the exact engine's tests
    assert tc(8, 7, 7, 3, 5, 3, 1, 0)[0] == 0
    assert tc(8, 32, 32, 48, 64, 3, 1, 1)[0] == 0
    assert L.livae_tc_conv5pool_supported(8, 64, 64, 3, 32) == 0
    # scratch sizes the Python side allocates from are positive and stable across calls
    assert L.livae_elbo_scratch_floats() == L.livae_elbo_scratch_floats() > 0
    assert L.livae_l2norm_scratch_floats() > 0 and L.livae_ssim_box_ws_floats(2048, 128) > 0


def _call_with(L, _lib, name, size, tc_desc):
    import ctypes as C
    args, keep = [], []
    for c in _lib._SIGS[name]:
        if c in "ps":
            args.append(None)
        elif c in "il":
            args.append(size)
        elif c == "f":
            args.append(1.0)
        elif c == "t":
            keep.append(_lib.TcConvDesc(*tc_desc)); args.append(C.byref(keep[-1]))
        else:
            keep.append(_lib.ConvDesc()); args.append(C.byref(keep[-1]))
    rc = getattr(L, name)(*args)
    return rc, (L.livae_last_error().decode(errors="replace") if rc else "")


def test_argument_errors_and_empty_problems_at_the_c_abi(libpath):
    """Error behaviour of every int-returning entry point, checked without a device: argument validation runs before any
    CUDA call, so (a) NULL device pointers with non-empty sizes are refused with rc < 0 and a message that names the
    function, never launched; (b) an empty problem (all sizes 0) is either a no-op (rc 0) or an argument error -- never a
    CUDA call (rc > 0 is a cudaError_t); nothing crashes."""
    from livae import _lib
    L = _lib.lib()
    noop = []

    def names_it(msg, name):
        # "<launcher>: what is wrong"; the ROI gather shares the sub-pixel gather's launcher and reports under its name
        who = msg.split(":")[0]
        return bool(msg) and (who in name or "_".join(who.split("_")[:2]) in name)

    for name in _lib._SIGS:
        rc, msg = _call_with(L, _lib, name, 0, (0,) * 11)
        assert rc <= 0, (name, rc, msg)
        if rc == 0:
            noop.append(name)
        else:
            assert names_it(msg, name), (name, msg)
        rc, msg = _call_with(L, _lib, name, 64, (8, 32, 32, 64, 64, 3, 3, 1, 1, 0, 0))
        assert rc < 0 and names_it(msg, name), (name, rc, msg)
    # the tensor-core entry points treat an empty batch as nothing to do
    assert {"livae_tc_conv", "livae_tc_conv_dgrad", "livae_tc_conv_wgrad", "livae_upfold_fwd"} <= set(noop)


def test_the_stub_printed_in_integration_md_binds(libpath):
    """the reference-side ctypes stub INTEGRATION.md shows a maintainer is executed as written (library path substituted):
    it binds, its argument list has the header's arity, and livae_last_error comes back as bytes"""
    import ctypes
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = re.search(r"```python\nimport ctypes, torch\n(.*?)```", doc, re.S).group(1)
    assert 'ctypes.CDLL("liblivae_sm100.so")' in code
    ns = {}
    exec("import ctypes, torch\n" + code.replace('ctypes.CDLL("liblivae_sm100.so")', f"ctypes.CDLL({libpath!r})"), ns)
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "livae_b200.h")).read(), flags=re.S)
    params = re.search(r"\blivae_rot_sample_fwd\s*\((.*?)\)\s*;", hdr, re.S).group(1)
    assert len(ns["_lib"].livae_rot_sample_fwd.argtypes) == len([p for p in params.split(",") if p.strip()])
    assert isinstance(ns["_lib"].livae_last_error(), bytes) and callable(ns["rotate"])
