"""CPU, world_size 2, gloo: the host-side data-parallel logic (site sharding, gradient averaging)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "li-vae_b200")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from livae.parallel import GradAverager, shard_sites
    sites = torch.stack([torch.zeros(101, dtype=torch.int32), torch.arange(101, dtype=torch.int32),
                         torch.arange(101, dtype=torch.int32) * 2], 1)
    mine = shard_sites(sites, rank, world, seed=5)
    flat = torch.full((1000,), float(rank + 1))
    flat[rank] = 10.0
    GradAverager(flat)()
    gathered = [None] * world
    dist.all_gather_object(gathered, mine[:, 1].tolist())
    if rank == 0:
        torch.save({"flat": flat, "shards": gathered}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_grad_average(tmp_path):
    out = str(tmp_path / "r.pt")
    mp.spawn(_worker, args=(2, 29611, out), nprocs=2, join=True)
    r = torch.load(out)
    a, b = r["shards"]
    assert len(a) == len(b) == 50 and not set(a) & set(b) and len(set(a) | set(b)) == 100
    flat = r["flat"]
    assert torch.allclose(flat[2:], torch.full((998,), 1.5))
    assert flat[0].item() == (10.0 + 2.0) / 2 and flat[1].item() == (1.0 + 10.0) / 2
