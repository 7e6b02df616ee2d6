"""C4 (BASELINE.json configs[3]): STN pre-training inner step (reference scripts/pretrain_stn.py:93-120) on synthetic
128x128 patches, batch 8192: two model.encoder() passes (x and its rotated copy), cycle-consistency loss, backward
(reaches the STN localisation only), clip 5.0, AdamW over the STN parameters.  Secondary number; bench.py's
headline stays C3.  usage: python tools/bench_stn.py [batch]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "li-vae_b200")]
import numpy as np
import torch
import livae
from livae import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
P = 128
dev = torch.device("cuda")
livae.set_engine("tc")
torch.manual_seed(1234)
m = livae.RVAE(latent_dim=2, in_channels=1, patch_size=P).to(dev)
stn_params = list(m.encoder.rotation_stn.parameters())
opt = torch.optim.AdamW(stn_params, lr=1e-3, weight_decay=1e-5)
g = torch.Generator(device="cpu").manual_seed(7)
batches = []
for _ in range(2):                      # 2 x (2 x 537 MB), cycled: larger than L2
    x = torch.rand(B, 1, P, P, generator=g).to(dev)
    ang = (torch.rand(B, generator=g) * 2 * np.pi).to(dev)
    xr = ops.rot_sample(x, ops.angle_to_cs(ang), 1.0)
    batches.append((x, xr, ang))


def step(b):
    x, xr, ang = b
    opt.zero_grad(set_to_none=True)
    _, _, th0 = m.encoder(x)
    _, _, th1 = m.encoder(xr)
    loss = livae.cycle_consistency_loss(th0, th1, ang)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(stn_params, max_norm=5.0)
    opt.step()
    return loss


for i in range(3):
    l0 = step(batches[i % 2])
torch.cuda.synchronize()
N = 6
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(N):
    l1 = step(batches[i % 2])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / N
roof = max(1.105e9 * B / 1376.8e12, 8.8e6 * B / 6551e9) * 1e3
print(json.dumps({"workload": "C4: STN pre-training inner step, P=128, incl. clip 5.0 + AdamW(STN)", "batch": B,
                  "ms_per_step": ms, "patches_per_s": B / ms * 1e3, "loss_first": float(l0.detach()), "loss_last": float(l1.detach()),
                  "step_roofline_ms": roof, "frac_of_roofline": roof / ms,
                  "accounting": "1.105 GFLOP and 8.8 MB fp32 boundary bytes per patch incl. the encoder convs whose outputs the loss never reads (SURVEY 8d)"}))
