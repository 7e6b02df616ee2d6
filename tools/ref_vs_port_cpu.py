"""Build container only: how representative is bench.py's CPU arm (the oracle PORT of the step) of the UNMODIFIED
reference trainer?  /root/reference cannot travel to the GPU box, so `bench.py --impl reference` times
oracle/rvae.py there; this script times, on the same host cores and the same batches,

  ref   the reference's own `train_rvae_one_epoch` (/root/reference/src/livae/train.py:286-445: forward, loss, backward,
        clip, AdamW AND its per-step metric block) on its own `RVAE` modules, --no-amp path, and
  port  bench.cpu_reference_step_rate (oracle.rvae.rvae_full_step + clip + AdamW, no metric block),

each in its own process (the two packages share the module name `livae`).  Output: one JSON line; the committed copy is
profiles/r02_cpu_reference_vs_port.json.

    python tools/ref_vs_port_cpu.py [steps] [batch]
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P, LATENT = 128, 2


def run_ref(steps, batch):
    sys.path.insert(0, ROOT)
    import numpy as np  # noqa: F401
    import torch
    from oracle import ref_loader
    from oracle import rvae as O
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    ref_loader.load()
    from livae.loss import RVAELoss
    from livae.model import RVAE
    from livae.train import MetricLogger, train_rvae_one_epoch
    params = O.make_params(O.rvae_param_shapes(P, LATENT), seed=1234, stn_head_std=0.5)
    batchdata = O.make_lattice_batch(batch, P, seed=2024)
    model = RVAE(latent_dim=LATENT, in_channels=1, patch_size=P)
    model.load_state_dict(params, strict=True)
    crit = RVAELoss(beta=10.0, gamma=10.0)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)      # scripts/train_rvae.py:157-159
    dev = torch.device("cpu")

    def epoch(n):
        train_rvae_one_epoch(model, [batchdata] * n, opt, crit, MetricLogger(), dev,
                             canonical_weight=0.2, scaler=None, grad_max_norm=20.0)

    epoch(1)
    t0 = time.perf_counter()
    epoch(steps)
    dt = time.perf_counter() - t0
    print(json.dumps({"patches_per_s": batch * steps / dt, "ms_per_step": dt / steps * 1e3,
                      "threads": torch.get_num_threads(), "torch": torch.__version__}))


def run_port(steps, batch):
    sys.path.insert(0, ROOT)
    import torch
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    import bench
    rate, ms, _, _ = bench.cpu_reference_step_rate(steps, 1, sample_b=batch)
    print(json.dumps({"patches_per_s": rate, "ms_per_step": ms, "threads": torch.get_num_threads(),
                      "torch": torch.__version__}))


def main():
    if len(sys.argv) > 1 and sys.argv[1] in ("--ref", "--port"):
        (run_ref if sys.argv[1] == "--ref" else run_port)(int(sys.argv[2]), int(sys.argv[3]))
        return
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    out = {"what": "CPU step rate of the unmodified reference trainer vs the oracle port bench.py times on the GPU box",
           "config": {"patch_size": P, "latent_dim": LATENT, "batch": batch, "steps": steps,
                      "step": "model(x) + encoder(x_rot) + RVAELoss(beta=10,gamma=10) + 0.2*canonical MSE + backward + "
                              "clip 20 + AdamW; the reference arm also runs train.py's per-step metric block"}}
    for arm in ("ref", "port", "ref", "port"):              # interleaved, best of two
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--" + arm, str(steps), str(batch)],
                           capture_output=True, text=True, check=True)
        d = json.loads(r.stdout.strip().splitlines()[-1])
        if arm not in out or d["patches_per_s"] > out[arm]["patches_per_s"]:
            out[arm] = d
    out["port_over_ref"] = out["port"]["patches_per_s"] / out["ref"]["patches_per_s"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
