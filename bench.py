"""bench.py -- rVAE train patches/s (fwd + bwd + ELBO + clip + AdamW) on B200.

Workload (BASELINE.json configs[2], "C3" in SURVEY.md section 8): rVAE latent_dim=2 on synthetic
128x128 lattice patches, batch 2048 per GPU, FULL reference step (train.py:373-397): model(x) +
model.encoder(x_rot) + RVAELoss(beta=10, gamma=10, cycle) + 0.2*canonical MSE + backward +
clip_grad_norm(20) + AdamW.  One "step" = one batch.  N > 1: one process per GPU (torchrun), each
rank trains on its own shard of the patches (weak scaling), gradients averaged with one NCCL
all-reduce of the flat gradient buffer.

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU arm (oracle port)

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "li-vae_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "rvae_train_patches_per_sec"
UNIT = "patches/s"
P, LATENT = 128, 2
BETA, GAMMA, CANON_W, MAX_NORM = 10.0, 10.0, 0.2, 20.0
# algorithmic work of the FULL step per 128x128 patch
FLOP_PER_PATCH = 2.90e9
BYTES_PER_PATCH_FP32 = 17.8e6


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's torch-CPU port of the reference step
# ------------------------------------------------------------------------------------------
def cpu_reference_step_rate(steps, warmup, sample_b=32):
    """Times the reference algorithm's CPU path (oracle/rvae.py restatement: the same ATen CPU ops
    the reference's nn.Modules dispatch to) on a bounded sample of the workload: batches of
    `sample_b` 128x128 patches.  /root/reference does not exist on the GPU box, so the port is used."""
    from oracle import rvae as O
    torch.manual_seed(0)
    params = O.make_params(O.rvae_param_shapes(P, LATENT), seed=1234, stn_head_std=0.5)
    x, xr, ang = O.make_lattice_batch(sample_b, P, seed=2024)
    eps = torch.from_numpy(np.random.default_rng(99).standard_normal((sample_b, LATENT))).float()
    m = {k: torch.zeros_like(v) for k, v in params.items()}
    v = {k: torch.zeros_like(v_) for k, v_ in params.items()}

    def one(t):
        _, grads = O.rvae_full_step(params, x, xr, ang, eps, beta=BETA, gamma=GAMMA, canonical_weight=CANON_W)
        tot = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
        coef = torch.clamp(MAX_NORM / (tot + 1e-6), max=1.0)
        for k in params:                      # AdamW (scripts/train_rvae.py:157-159)
            g = grads[k] * coef
            params[k].mul_(1 - 1e-3 * 1e-5)
            m[k].lerp_(g, 0.1)
            v[k].mul_(0.999).addcmul_(g, g, value=0.001)
            params[k].addcdiv_(m[k] / (1 - 0.9 ** t), (v[k] / (1 - 0.999 ** t)).sqrt_().add_(1e-8), value=-1e-3)

    for i in range(warmup):
        one(i + 1)
    t0 = time.perf_counter()
    for i in range(steps):
        one(warmup + i + 1)
    dt = time.perf_counter() - t0
    return sample_b * steps / dt, dt / steps * 1e3, sample_b


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 8))
    warm = max(1, min(args.warmup, 2))
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1 to every rank)
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:
        torch.set_num_threads(max(1, os.cpu_count() or 1))
    rate, ms, sb = cpu_reference_step_rate(steps, warm)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{steps} steps of batch {sb} (128x128 patches) of the same FULL rVAE step, "
                                   f"torch {torch.__version__} CPU fp32, {cores} threads of {os.cpu_count()} cpus"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, world):
    return {"workload": "C3: rVAE latent_dim=2, synthetic 128x128 lattice patches, FULL train step "
                        "(model(x) + encoder(x_rot) + RVAELoss(beta=10,gamma=10,cycle) + 0.2*canonical MSE + "
                        "backward + clip 20 + AdamW)",
            "patch_size": P, "latent_dim": LATENT, "batch_per_gpu": args.batch, "global_batch": args.batch * world,
            "parallelism": f"dp{world}", "l2_policy": f"{args.nbatches} distinct resident batches of "
            f"{2 * args.batch * P * P * 4 / 1e6:.0f} MB (x + x_rot) cycled, each larger than the 126 MB L2",
            "x_rot": "rot_sample(x, angle~U(0,2pi)) with reflection padding (synthetic pair)",
            "engine": args.engine}


# ------------------------------------------------------------------------------------------
# synthetic HAADF-like data, generated on the device outside the timed region
# ------------------------------------------------------------------------------------------
def synth_haadf(hw, a, angle_deg, seed, device):
    g = torch.Generator(device=device).manual_seed(seed)
    r = torch.arange(hw, device=device, dtype=torch.float32)
    yy, xx = torch.meshgrid(r, r, indexing="ij")
    img = torch.zeros((hw, hw), device=device)
    t0 = np.deg2rad(angle_deg)
    for k in range(3):
        t = t0 + k * np.pi / 3
        img += torch.cos((2 * np.pi / a) * (np.cos(t) * xx + np.sin(t) * yy) + 1.3 * k)
    img += 0.15 * torch.randn((hw, hw), device=device, generator=g)
    img -= img.min()
    img /= img.max()
    return img


def make_batches(args, device, rank):
    from livae import ops
    n_img, hw = 2, 2048
    imgs = torch.stack([synth_haadf(hw, 12.0, 3.75 * (k + 2 * rank), 1000 + k + 16 * rank, device)
                        for k in range(n_img)]).contiguous()
    g = torch.Generator(device="cpu").manual_seed(2024 + rank)
    batches = []
    for _ in range(args.nbatches):
        sites = torch.stack([torch.randint(0, n_img, (args.batch,), generator=g),
                             torch.randint(96, hw - 96, (args.batch,), generator=g),
                             torch.randint(96, hw - 96, (args.batch,), generator=g)], 1).to(torch.int32)
        x = ops.patch_minmax_(ops.patch_gather(imgs, sites.to(device), P))
        ang = (torch.rand(args.batch, generator=g) * 2 * np.pi).to(device)
        xr = ops.rot_sample(x, ops.angle_to_cs(ang), 1.0)
        batches.append((x, xr, ang))
    return batches


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev_index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(dev_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, val in zip(names, f[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# per-kernel-family accounting for the roofline object
# ------------------------------------------------------------------------------------------
def describe_call(name, a):
    """-> (family key, algorithmic flops, algorithmic bytes) of one C-ABI call"""
    nb = lambda t: 0 if t is None else t.numel() * t.element_size()
    if name in ("livae_conv_fwd", "livae_conv_bwd"):
        d = a[0]._obj
        if d.kind == 0:
            ho = (d.Hin + 2 * d.pad - d.kh) // d.stride + 1
            wo = (d.Win + 2 * d.pad - d.kw) // d.stride + 1
            macs = d.B * ho * wo * d.Cout * d.kh * d.kw * d.Cin
        else:
            macs = d.B * d.Hin * d.Win * d.Cin * d.Cout * d.kh * d.kw
        key = f"{name[6:]}[{'convT' if d.kind else 'conv'} {d.Cin}->{d.Cout} k{d.kh} s{d.stride} {d.Hin}x{d.Win}]"
        if name == "livae_conv_fwd":
            return key, 2.0 * macs, nb(a[1]) + nb(a[4])
        n = (a[6] is not None) + (a[8] is not None)      # wgrad, dgrad
        return key, 2.0 * macs * n, nb(a[1]) + nb(a[4]) + nb(a[8])
    by = sum(nb(t) for t in a if isinstance(t, torch.Tensor))
    if name in ("livae_tc_conv", "livae_tc_conv_dgrad", "livae_tc_conv_wgrad"):
        d = a[0]._obj
        ho = (d.Hin + 2 * d.pad - d.kh) // d.stride + 1
        wo = (d.Win + 2 * d.pad - d.kw) // d.stride + 1
        macs = d.B * ho * wo * d.Cout * d.kh * d.kw * d.Cin
        key = f"{name[6:]}[{d.Cin}->{d.Cout} k{d.kh} s{d.stride} {d.Hin}x{d.Win}]"
        return key, 2.0 * macs, by
    if name.startswith("livae_thin_conv"):
        ints = [v for v in a if isinstance(v, int)]
        return f"{name[6:]}[{','.join(str(v) for v in ints[:4])}]", 0.0, by
    return name[6:], 0.0, by


def profile_families(step_fn, batches, nsteps=2):
    from livae import _lib
    _lib.PROFILE = []
    for i in range(nsteps):
        step_fn(batches[i % len(batches)])
    torch.cuda.synchronize()
    rec = _lib.PROFILE
    _lib.PROFILE = None
    fam = {}
    for name, a, e0, e1 in rec:
        key, fl, by = describe_call(name, a)
        f = fam.setdefault(key, {"ms": 0.0, "calls": 0, "flops": 0.0, "bytes": 0.0})
        f["ms"] += e0.elapsed_time(e1); f["calls"] += 1; f["flops"] += fl; f["bytes"] += by
    return fam, len(rec) / nsteps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--nbatches", type=int, default=3)
    ap.add_argument("--engine", default="tc", help="convolution engine: tc (tcgen05, bf16 storage) | f32 (exact SIMT)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    import livae
    from livae import _lib, optim
    from livae.train import train_rvae_step, DevicePrefetcher
    livae.set_engine(args.engine)

    torch.manual_seed(1234)
    model = livae.RVAE(latent_dim=LATENT, in_channels=1, patch_size=P).to(device)
    if world > 1:
        for p_ in model.parameters():
            dist.broadcast(p_.data, 0)
    crit = livae.RVAELoss(beta=BETA, gamma=GAMMA)
    opt = optim.FlatAdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
    torch.manual_seed(99 + rank)           # eps stream

    reduce_grads = None
    if world > 1:
        from livae.parallel import GradAverager
        reduce_grads = GradAverager(opt.flat_grad)   # ONE NCCL all-reduce of the 9 MB flat gradient per step

    batches = make_batches(args, device, rank)
    host = [tuple(t.cpu().pin_memory() for t in b) for b in batches]

    def step(batch):
        return train_rvae_step(model, opt, crit, batch, device, CANON_W, MAX_NORM, reduce_grads)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(batch_src, read_loss):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        last = None
        feed = (batch_src[i % len(batch_src)] for i in range(args.steps))
        if read_loss:
            # e2e: host batches through the trainer's own prefetcher (livae.train.DevicePrefetcher, the loop
            # train_rvae_one_epoch runs): batch i+1 is copied from pinned memory while step i computes
            prefetch.loader = feed
            feed = prefetch
        for batch in feed:
            out = step(batch)
            if read_loss:
                last = out[1].item()        # device -> host read of the step's loss, every step
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, last

    for i in range(args.warmup):
        out = step(batches[i % len(batches)])
        barrier()      # also warms the NCCL barrier itself: its first call cost ~50 ms inside a 6-step timed region at N=2
    loss0 = out[1].item()
    # warm-up of the e2e path too: the prefetcher's two staging slots (2 x 268 MB) are allocated on first use, and
    # that first cudaMalloc cost 10-110 ms inside a 6-step timed region
    prefetch = DevicePrefetcher((host[i % len(host)] for i in range(2)), device)
    for b in prefetch:
        step(b)[1].item()
    if not np.isfinite(loss0):
        raise SystemExit(f"non-finite loss {loss0}")

    L = _lib.lib()
    has_counter = hasattr(L, "livae_launch_count")
    if has_counter:
        L.livae_launch_count.restype = __import__("ctypes").c_int64
        c0 = L.livae_launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    ms, _ = timed(batches, read_loss=False)
    launches = (L.livae_launch_count() - c0) if has_counter else None
    ms_e2e, last_loss = timed(host, read_loss=True)
    clocks = sampler.stop() if sampler else None

    value = world * args.batch * args.steps / (ms / 1e3)
    e2e = world * args.batch * args.steps / (ms_e2e / 1e3)

    if rank == 0:
        pk = peaks()
        # per-kernel accounting runs on rank 0 alone: no collective inside it (the other ranks wait at the barrier)
        local_step = lambda batch: train_rvae_step(model, opt, crit, batch, device, CANON_W, MAX_NORM, None)
        fam, calls_per_step = profile_families(local_step, batches)
        tot = sum(f["ms"] for f in fam.values())
        top_key, top = max(fam.items(), key=lambda kv: kv[1]["ms"])
        per_launch_ms = top["ms"] / top["calls"]
        if top["flops"] > 0:
            ach = top["flops"] / top["calls"] / (per_launch_ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": top_key, "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                    "frac": ach / pk["tf_sust"], "traffic": None, "peak_source": pk["src"] + " bf16 sustained",
                    "share_of_step": top["ms"] / tot, "avg_launch_ms": per_launch_ms}
        else:
            ach = top["bytes"] / top["calls"] / (per_launch_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": top_key, "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
                    "frac": ach / pk["hbm"], "traffic": None, "peak_source": pk["src"],
                    "share_of_step": top["ms"] / tot, "avg_launch_ms": per_launch_ms}
        # measured DRAM traffic of that kernel (ncu --set full capture committed under profiles/), else null
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(top_key)
            if tr:
                roof["traffic"] = tr["bytes_per_launch"]
                roof["traffic_source"] = tr["source"]
                roof["algorithmic_bytes_per_launch"] = top["bytes"] / top["calls"]
        except Exception:
            pass
        t_step = ms / args.steps / 1e3
        t_roof = max(FLOP_PER_PATCH * args.batch / (pk["tf_sust"] * 1e12),
                     BYTES_PER_PATCH_FP32 * args.batch / (pk["hbm"] * 1e9))
        kernels = sorted(({"kernel": k, "ms_per_step": f["ms"] / 2, "calls_per_step": f["calls"] / 2,
                           "share": f["ms"] / tot,
                           "tflops": (f["flops"] / (f["ms"] * 1e-3) / 1e12) if f["flops"] else None,
                           "gbs": f["bytes"] / (f["ms"] * 1e-3) / 1e9} for k, f in fam.items()),
                         key=lambda r: -r["ms_per_step"])[:40]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.engine == "f32" else "bf16", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": world * sum(t.numel() * t.element_size() for t in host[0]),
                    "d2h_bytes_per_step": world * 4},
            "gpu_launches": int(launches) if launches is not None else int(calls_per_step * args.steps),
            "clocks": clocks, "roofline": roof,
            "step_roofline": {"t_roof_ms": t_roof * 1e3, "t_step_ms": t_step * 1e3, "frac": t_roof / t_step,
                              "accounting": "FULL step, 2.90 GFLOP/patch vs bf16 sustained peak; 17.8 MB/patch "
                                            "fp32 layer-boundary bytes vs measured HBM (SURVEY.md 8d)"},
            "kernels": kernels, "final_loss": last_loss,
        }
        if world == 1 and not args.no_cpu_baseline:
            rate, cms, sb = cpu_reference_step_rate(3, 1)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"3 steps of batch {sb} of the same FULL step, oracle torch-CPU port, "
                                              f"{cms:.0f} ms/step"}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
