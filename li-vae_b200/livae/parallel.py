"""Data-parallel plumbing for the rVAE step: patches are independent, so
the path shards by SITE with no data-path collective; the only exchange is ONE all-reduce (NCCL
over NVLink on the GPU box) of the flat fp32 gradient buffer per step, followed by global-norm
clipping and the optimiser update on the reduced gradients.  The reference has no distributed
code; the correctness contract is "N ranks x local batch B/N == 1 rank x batch B".
"""
from __future__ import annotations

import torch

__all__ = ["shard_sites", "GradAverager"]


def shard_sites(sites: torch.Tensor, rank: int, world: int, seed: int = 0, drop_remainder: bool = True):
    """Seeded global permutation of the (img, cy, cx) site table, then rank::world striding, so every
    rank sees a disjoint, equally sized, statistically identical shard (all ranks take equal steps)."""
    n = sites.shape[0]
    g = torch.Generator(device="cpu").manual_seed(seed)
    perm = torch.randperm(n, generator=g)
    if drop_remainder:
        perm = perm[: n - n % world]
    return sites[perm[rank::world].to(sites.device)]


class GradAverager:
    """all-reduce(sum) of a flat gradient buffer, then 1/world -- called between backward and clip"""

    def __init__(self, flat_grad: torch.Tensor, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.flat = flat_grad
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def __call__(self):
        if self.world == 1:
            return
        self.dist.all_reduce(self.flat, op=self.dist.ReduceOp.SUM, group=self.group)
        self.flat.mul_(1.0 / self.world)
