"""CPU: the drop-in surface the reference's scripts import, the host-side pre-processing and site finding
against vectors produced by the reference (tests/golden/make_golden_r2.py), and the DataLoader-worker recipe
mechanism of livae.data (workers produce recipes; the pixels are made on the GPU in the main process)."""
import os
import sys
import random

import numpy as np
import pytest
import torch

from tests.golden.make_golden_r2 import synth_lattice
from tests.util import load_golden

# The import statements of the reference's three training scripts, verbatim:
# scripts/train_rvae.py:14-24, scripts/train_vae.py:14-24, scripts/pretrain_stn.py:13-16.
SCRIPT_IMPORTS = {
    "scripts/train_rvae.py": """
from livae.data import PairedAdaptiveLatticeDataset
from livae.loss import RVAELoss
from livae.model import RVAE
from livae.train import (
    MetricLogger,
    evaluate_rvae,
    log_reconstructions_tensorboard,
    log_scalar_metrics_tensorboard,
    train_rvae_one_epoch,
)
from livae.utils import load_image_from_h5
""",
    "scripts/train_vae.py": """
from livae.data import AdaptiveLatticeDataset, default_transform
from livae.loss import VAELoss
from livae.model import VAE
from livae.train import (
    MetricLogger,
    evaluate,
    log_reconstructions_tensorboard,
    log_scalar_metrics_tensorboard,
    train_one_epoch,
)
from livae.utils import load_image_from_h5
""",
    "scripts/pretrain_stn.py": """
from livae.data import PairedAdaptiveLatticeDataset
from livae.loss import cycle_consistency_loss
from livae.model import RVAE
from livae.utils import load_image_from_h5
""",
}
REF = "/root/reference"


@pytest.mark.parametrize("script", sorted(SCRIPT_IMPORTS))
def test_script_import_lines_resolve(script):
    ns = {}
    exec(compile(SCRIPT_IMPORTS[script], script, "exec"), ns)
    import livae
    for name, obj in ns.items():
        if name.startswith("__"):
            continue
        assert obj.__module__.startswith("livae"), (name, obj.__module__)
    assert os.path.dirname(livae.__file__).endswith(os.path.join("li-vae_b200", "livae"))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("script", sorted(SCRIPT_IMPORTS))
def test_script_import_lists_are_current(script):
    """the committed lists above are the scripts' actual `from livae...` statements"""
    src = open(os.path.join(REF, script)).read()
    head = src[:src.index("\ndef ")]
    want = "".join(ln + "\n" for ln in _livae_imports(head))
    assert want.strip() == SCRIPT_IMPORTS[script].strip()
    exec(compile(head, script, "exec"), {})          # the whole header, third-party imports included


def _livae_imports(head):
    out, keep = [], False
    for ln in head.splitlines():
        if ln.startswith("from livae"):
            keep = True
        if keep:
            out.append(ln)
            if ")" in ln or (ln.startswith("from livae") and "(" not in ln):
                keep = False
    return out


def test_package_root_exports_reference_names():
    import livae
    # src/livae/__init__.py:36-70 of the reference
    for name in ("PatchDataset", "default_transform", "normalize_image", "bandpass_filter", "fft_spectra",
                 "lowpass_filter", "highpass_filter", "VAELoss", "VAE", "RVAE", "Encoder", "Decoder", "RotationSTN",
                 "train_one_epoch", "evaluate", "evaluate_rotation_invariance", "log_reconstructions_tensorboard",
                 "log_scalar_metrics_tensorboard", "MetricLogger", "compute_psnr", "compute_ssim",
                 "compute_reconstruction_metrics", "compute_latent_metrics", "compute_atom_detection_metrics",
                 "compute_all_metrics", "load_image_from_h5", "estimate_lattice_constant"):
        assert hasattr(livae, name), name


def test_filters_and_lattice_constant_match_reference():
    from livae import filter as F
    from livae.utils import estimate_lattice_constant
    g = load_golden("preproc.npz")
    img = synth_lattice(256, 14.0, 11.0, 5)
    rect = np.random.default_rng(6).random((96, 128)) * 1000.0
    samp = lambda a: a[::7, ::5]
    np.testing.assert_allclose(samp(F.bandpass_filter(img, 10, 60)), g["band"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(samp(F.lowpass_filter(rect, 20)), g["low"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(samp(F.highpass_filter(rect, 7)), g["high"], rtol=0, atol=1e-9)
    np.testing.assert_array_equal(samp(F.normalize_image(rect)), g["norm"])
    mag, ph = F.fft_spectra(rect)
    np.testing.assert_allclose(samp(mag), g["mag"], rtol=1e-12)
    np.testing.assert_allclose(samp(ph), g["phase"], rtol=0, atol=1e-9)
    for k, (hw, a) in enumerate(((256, 14.0), (512, 16.0), (512, 19.0), (384, 24.0))):
        assert estimate_lattice_constant(synth_lattice(hw, a, 7.0 + 13 * k, 20 + k)) == float(g[f"lattice_{k}"])
    assert estimate_lattice_constant(np.random.default_rng(1).random((128, 128))) == float(g["lattice_noise"])
    with pytest.raises(ValueError):
        F.bandpass_filter(rect, 30, 10)
    with pytest.raises(ValueError):
        F.lowpass_filter(np.zeros((2, 3, 4)), 3)
    assert np.all(F.normalize_image(np.full((4, 4), 3.0)) == 0)


def test_site_finding_matches_reference_loops():
    """vectorised adaptive lattice sites == the reference's per-atom loops + union-find on the same peaks"""
    from livae.data import AdaptiveLatticeDataset, PatchDataset, generate_lattice_grid
    g = load_golden("sites.npz")
    imgs = [synth_lattice(512, 16.0, 7.0, 1), synth_lattice(512, 19.0, 33.0, 2)]
    ad = AdaptiveLatticeDataset(imgs, patch_size=64, padding=16, transform=None)
    pd = PatchDataset(imgs, patch_size=64, padding=8, transform=None)
    for i in range(2):
        assert ad.sample_coords[i].shape == g[f"sample_coords_{i}"].shape
        np.testing.assert_allclose(ad.sample_coords[i], g[f"sample_coords_{i}"], rtol=0, atol=1e-9)
        np.testing.assert_array_equal(ad.labels[i], g[f"labels_{i}"])
        np.testing.assert_array_equal(pd.atom_coords[i], g[f"atom_coords_{i}"])
        assert abs(ad.images[i].sum() - float(g[f"image_sum_{i}"])) < 1e-6
    assert len(ad) == sum(len(g[f"sample_coords_{i}"]) for i in range(2))
    np.testing.assert_allclose(generate_lattice_grid((100, 120), 11.3, (2.5, 1.0)), g["grid"], rtol=0, atol=1e-12)
    with pytest.raises(IndexError):
        ad._check(len(ad))


def test_peak_local_max_semantics():
    from livae.sites import get_clean_peaks, peak_local_max
    img = np.zeros((40, 40))
    img[10, 10] = 1.0; img[10, 13] = 0.9; img[25, 30] = 0.5; img[1, 20] = 2.0; img[30, 5] = 0.004
    # (10,13) is within min_distance=4 of the brighter (10,10); (1,20) is in the excluded border; (30,5) < 1 % of max
    pk = peak_local_max(img, min_distance=4, threshold_rel=0.01)
    assert pk.tolist() == [[10, 10], [25, 30]]
    assert peak_local_max(np.zeros((8, 8)), min_distance=2).shape == (0, 2)
    # plateau: two equal neighbours -> one peak
    img2 = np.zeros((20, 20)); img2[8, 8] = img2[8, 9] = 1.0
    assert len(peak_local_max(img2, min_distance=3)) == 1
    # refinement moves a peak to the 5x5 argmax (first in row-major order on ties)
    assert get_clean_peaks(img, min_distance=4).tolist() == [[10, 10], [25, 30]]


def _peak_local_max_loops(img, min_distance, threshold_rel=None):
    """scikit-image 0.25.2 (the reference's pinned version, uv.lock) `peak_local_max(image, min_distance, threshold_rel)`
    with its defaults, written out as loops from its published algorithm -- `_get_peak_mask` (pixel equals the maximum of
    its (2d+1)^2 window, edges replicated; strictly above max(image.min(), threshold_rel * image.max())),
    `_exclude_border` (d pixels on every side), `_get_high_intensity_peaks` (stable sort by falling intensity) and
    `ensure_spacing` (walk the sorted list; an accepted peak rejects every other candidate at Chebyshev distance < d)."""
    h, w = img.shape
    d = int(min_distance)
    thr = img.min()
    if threshold_rel is not None:
        thr = max(thr, threshold_rel * img.max())
    cand = []
    for y in range(d, h - d):
        for x in range(d, w - d):
            win = img[max(y - d, 0):y + d + 1, max(x - d, 0):x + d + 1]      # inside the border the window is never clipped
            if img[y, x] == win.max() and img[y, x] > thr:
                cand.append((y, x))
    cand.sort(key=lambda p: -img[p])                                        # list.sort is stable: row-major among equals
    rejected, out = set(), []
    for i, p in enumerate(cand):
        if i in rejected:
            continue
        out.append(p)
        for j, q in enumerate(cand):
            if j != i and max(abs(p[0] - q[0]), abs(p[1] - q[1])) < d:
                rejected.add(j)
    return np.asarray(out, dtype=np.int64).reshape(-1, 2)


def test_peak_local_max_equals_the_written_out_algorithm():
    """the vectorised peak finder against the loop restatement above on noisy lattices, quantised images (plateaus and
    exact ties) and degenerate inputs; scikit-image itself is not installed, so this pins the implementation to a second,
    independently written statement of the documented algorithm rather than to the library"""
    from livae.sites import peak_local_max
    rng = np.random.default_rng(17)
    yy, xx = np.mgrid[:48, :56].astype(np.float64)
    lattice = np.cos(2 * np.pi * xx / 9.0) * np.cos(2 * np.pi * yy / 7.0)
    cases = [(lattice + rng.normal(0, 0.2, lattice.shape), 3, 0.01),
             (lattice + rng.normal(0, 0.05, lattice.shape), 2, 0.3),
             (np.round(rng.random((40, 40)) * 4) / 4, 2, 0.01),              # heavy ties and plateaus
             (np.round(rng.random((30, 33)) * 3), 1, None),
             (rng.random((25, 25)), 4, None),
             (rng.random((12, 12)), 5, 0.5),                                  # border wider than most of the image
             (np.full((9, 9), 2.5), 2, 0.01)]                                 # constant image: no peaks
    for img, d, tr in cases:
        got = peak_local_max(img, min_distance=d, threshold_rel=tr)
        want = _peak_local_max_loops(img, d, tr)
        assert got.shape == want.shape and np.array_equal(got, want), (img.shape, d, tr, len(got), len(want))
    assert len(_peak_local_max_loops(cases[0][0], 3, 0.01)) > 20


def _tiny_dataset(kind, transform):
    from livae import data as D
    rng = np.random.default_rng(3)
    imgs = [rng.random((128, 128)), rng.random((128, 128))]
    sites = [rng.uniform(40, 88, size=(7, 2)), rng.uniform(40, 88, size=(5, 2))]
    cls = {"paired": D.PairedAdaptiveLatticeDataset, "adaptive": D.AdaptiveLatticeDataset}[kind]
    return cls.from_sites(imgs, sites, patch_size=32, padding=8, transform=transform)


def test_worker_items_are_recipes_and_collate_into_a_batch():
    """DataLoader(num_workers=2) as the scripts build it (minus pin_memory: no GPU here): workers return recipes,
    torch's default collate packs them, nothing touches the CUDA library in the workers"""
    from torch.utils.data import DataLoader, random_split
    from livae import data as D
    ds = _tiny_dataset("paired", D.default_transform)
    assert len(ds) == 12
    train, val = random_split(ds, [10, 2], generator=torch.Generator().manual_seed(0))
    loader = DataLoader(train, batch_size=4, shuffle=True, num_workers=2, persistent_workers=True, prefetch_factor=1,
                        drop_last=True, generator=torch.Generator().manual_seed(1))
    batches = list(loader)
    assert len(batches) == 2
    seen = []
    for b in batches:
        assert isinstance(b, D.RecipeBatch) and len(b) == 4
        assert b.key == ds._key and b.indices.dtype == np.int64
        assert set(b.t) == {"scale", "angle", "flags", "shift"} and b.t["scale"].shape == (4,) and b.t["shift"].shape == (4, 2)
        assert b.angles.shape == (4,) and np.all((b.angles >= 0) & (b.angles <= 360))
        assert np.all((b.t["scale"] >= 0.9) & (b.t["scale"] <= 1.1))
        seen += b.indices.tolist()
    assert len(set(seen)) == 8 and set(seen) <= set(train.indices)
    del loader


def test_recipe_draws_follow_reference_order():
    """a recipe holds exactly the numbers the reference's item would have drawn from `random`, in its order:
    scale, hflip, vflip, shift_x, shift_y (default_transform, data.py:85-114), then the pair angle (data.py:695)"""
    from livae import data as D
    ds = _tiny_dataset("paired", D.default_transform)
    random.seed(42)
    t, ang = ds._draws(1)
    random.seed(42)
    scale = random.uniform(0.9, 1.1)
    h = random.random() < 0.5
    v = random.random() < 0.5
    sx = random.randint(-4, 4); sy = random.randint(-4, 4)
    a = random.uniform(0, 360)
    assert np.float32(scale) == t["scale"][0] and t["flags"][0] == (1 if h else 0) | (2 if v else 0)
    assert t["shift"][0].tolist() == [sy, sx] and ang[0] == a
    nt = _tiny_dataset("adaptive", None)
    assert nt._draws(3) == (None, None)


def test_dataset_pickles_without_device_state_and_has_no_cpu_path():
    import pickle
    from livae import data as D
    ds = _tiny_dataset("paired", None)
    clone = pickle.loads(pickle.dumps(ds))
    assert clone._src is None and len(clone) == len(ds)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            ds[0]                       # main-process items are made by the CUDA kernels: no CPU fallback


def test_utils_clean_state_dict_and_h5_error():
    from livae.utils import clean_state_dict, load_image_from_h5
    sd = clean_state_dict({"_orig_mod.encoder.fc_mu.weight": 1, "decoder.fc.bias": 2})
    assert sd == {"encoder.fc_mu.weight": 1, "decoder.fc.bias": 2}
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            load_image_from_h5("nope.h5")


def test_tensorboard_helpers_signature():
    from livae.train import log_scalar_metrics_tensorboard

    class W:
        def __init__(self):
            self.rows = []

        def add_scalar(self, *a):
            self.rows.append(a)

    w = W()
    log_scalar_metrics_tensorboard(w, {"a": 1.0, "b": 2.0}, 7, prefix="train/")
    assert w.rows == [("train/a", 1.0, 7), ("train/b", 2.0, 7)]


def _h5_expectations_hold(load, fake):
    for contents, name, want in fake.cases():
        fake.FILES["f.h5"] = contents
        if isinstance(want, type):
            with pytest.raises(want):
                load("f.h5", name)
        else:
            got = load("f.h5", name)
            assert got.shape == contents[want].shape and np.array_equal(got, contents[want]), (name, want)


def test_load_image_from_h5_dataset_selection(monkeypatch):
    """utils.py:111-185 of the reference: explicit path, base-name search, auto-detection (preferred names, then area,
    ties in visiting order), KeyError without a 2-D dataset -- through an in-memory stand-in for h5py (not installed here)"""
    from tests import fake_h5py
    monkeypatch.setitem(sys.modules, "h5py", fake_h5py.install())
    from livae.utils import load_image_from_h5
    _h5_expectations_hold(load_image_from_h5, fake_h5py)
    _h5_expectations_hold(lambda p, n: load_image_from_h5(__import__("pathlib").Path(p), n), fake_h5py)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_h5_expectations_are_the_references_behaviour():
    """the same table against the reference's own load_image_from_h5 (its process: the two packages share a name)"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from tests import fake_h5py as F; F.install()\n"
            "from oracle import ref_loader; ref_loader.load()\n"
            "from livae.utils import load_image_from_h5\n"
            "assert load_image_from_h5.__module__ == 'livae.utils' and '/root/reference' in sys.modules['livae'].__file__\n"
            "from tests.test_dropin_cpu import _h5_expectations_hold\n"
            "_h5_expectations_hold(load_image_from_h5, F); print('REF_OK')\n") % root
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd="/tmp")
    assert r.returncode == 0 and "REF_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


def test_dataset_plot_helpers_draw_the_sites_of_the_region(monkeypatch):
    """PatchDataset.plot_peaks / AdaptiveLatticeDataset.plot_lattice (reference data.py:252-289, 562-612) with a
    recording stand-in for matplotlib.pyplot (matplotlib is not installed here): the cropped image, the sites inside the
    region in region coordinates, atoms and empty sites as separate scatters"""
    import types
    from livae import data as D
    calls = []
    plt = types.ModuleType("matplotlib.pyplot")
    for name in ("figure", "imshow", "scatter", "axis", "show"):
        setattr(plt, name, lambda *a, _n=name, **k: calls.append((_n, a, k)))
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    monkeypatch.setitem(sys.modules, "matplotlib", mpl)
    monkeypatch.setitem(sys.modules, "matplotlib.pyplot", plt)

    img = np.arange(40.0 * 50).reshape(40, 50)
    sites = np.array([[5.0, 6.0], [12.5, 30.0], [20.0, 21.0], [35.0, 45.0]])
    ds = D.AdaptiveLatticeDataset.from_sites([img], [sites], patch_size=8, padding=2,
                                             labels=[np.array([1, 0, 1, 0])])
    ds.plot_lattice(0, size=20, offset=(4, 5))
    kinds = [c[0] for c in calls]
    assert kinds == ["figure", "imshow", "scatter", "axis", "show"]       # no empty site in this region: no second scatter
    assert calls[0][2] == {"figsize": (8, 8)} and np.array_equal(calls[1][1][0], img[4:24, 5:25])
    (xs, ys), kw = calls[2][1], calls[2][2]                      # atoms inside the region: (5,6) and (20,21)
    assert list(xs) == [1.0, 16.0] and list(ys) == [1.0, 16.0] and kw["s"] == 50 and kw["c"] == "red"
    calls.clear()
    ds.plot_lattice(0)                                           # whole image: two atoms, two empty sites
    assert [len(c[1][0]) for c in calls if c[0] == "scatter"] == [2, 2] and calls[1][1][0].shape == (40, 50)

    calls.clear()
    pd_ = D.PatchDataset.__new__(D.PatchDataset)                 # plot helper only: skip the site finding
    pd_.images, pd_.atom_coords = [img], [np.array([[5, 6], [30, 40]])]
    pd_.plot_peaks(0, size=16, offset=(0, 0))
    assert [c[0] for c in calls] == ["figure", "imshow", "scatter", "axis", "show"]
    assert calls[0][2] == {"figsize": (6, 6)} and list(calls[2][1][0]) == [6] and list(calls[2][1][1]) == [5]
    assert calls[2][2]["s"] == 30 and calls[3][1] == ("off",)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_public_surface_and_signatures_match_the_reference_modules():
    """every public function, class and public method (plus __init__ / forward / __getitem__) defined in the reference's
    seven modules exists here under the same name with the same positional parameter names in the same order; trailing
    extra parameters (train_rvae_one_epoch's reduce_grads) are allowed.  Read from the reference's SOURCE (ast): the two
    packages share the name `livae` and cannot be imported side by side."""
    import ast
    import importlib
    import inspect

    def ref_params(fn):
        return [a.arg for a in fn.args.posonlyargs + fn.args.args]

    def our_params(obj):
        ps = inspect.signature(obj).parameters.values()
        return [p.name for p in ps if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]

    problems = []
    for mod in ("data", "loss", "model", "train", "utils", "filter", "metrics"):
        tree = ast.parse(open(os.path.join(REF, "src", "livae", mod + ".py")).read())
        ours = importlib.import_module("livae." + mod)
        for node in tree.body:
            if isinstance(node, ast.FunctionDef) and not node.name.startswith("_"):
                obj = getattr(ours, node.name, None)
                if obj is None or our_params(obj)[:len(ref_params(node))] != ref_params(node):
                    problems.append((mod, node.name))
            elif isinstance(node, ast.ClassDef) and not node.name.startswith("_"):
                cls = getattr(ours, node.name, None)
                if cls is None:
                    problems.append((mod, node.name))
                    continue
                for f in node.body:
                    if isinstance(f, ast.FunctionDef) and (f.name in ("__init__", "forward", "__getitem__", "__len__")
                                                           or not f.name.startswith("_")):
                        m = getattr(cls, f.name, None)
                        if m is None or our_params(m)[:len(ref_params(f))] != ref_params(f):
                            problems.append((mod, node.name + "." + f.name))
    assert not problems, problems


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_every_livae_import_of_every_reference_script_resolves():
    """not only the three training scripts: each `from livae.<module> import <names>` anywhere in the reference's scripts/,
    top-level helper scripts and verify_*.py (Ray Tune drivers, comparison / t-SNE / visualisation scripts included) names
    something this package provides"""
    import ast
    import glob
    import importlib
    files = sorted(glob.glob(os.path.join(REF, "scripts", "*.py")) + glob.glob(os.path.join(REF, "*.py")))
    assert len(files) >= 10
    missing, seen = [], 0
    for path in files:
        for node in ast.walk(ast.parse(open(path).read())):
            if isinstance(node, ast.ImportFrom) and node.module and node.module.split(".")[0] == "livae":
                mod = importlib.import_module(node.module)
                for a in node.names:
                    seen += 1
                    if not hasattr(mod, a.name):
                        missing.append((os.path.basename(path), node.module, a.name))
    assert seen >= 40 and not missing, missing
