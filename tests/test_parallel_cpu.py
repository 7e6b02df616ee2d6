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


def test_shard_sites_partitions_the_table_at_every_world_size():
    """every site lands on exactly one rank, shards have equal length (all ranks take the same number of steps),
    the split depends on the seed only -- checked as index arithmetic on the C5-sized table (2 M sites) and on ragged /
    tiny / empty tables"""
    sys.path[:0] = [os.path.join(ROOT, "li-vae_b200")]
    from livae.parallel import shard_sites
    for n in (0, 1, 7, 101, 2_000_000):
        table = torch.stack([torch.arange(n) % 16, torch.arange(n), torch.arange(n) * 3], 1).to(torch.int32)
        for world in (1, 2, 3, 4, 8):
            shards = [shard_sites(table, r, world, seed=5) for r in range(world)]
            assert len({len(s) for s in shards}) == 1 and len(shards[0]) == n // world
            ids = torch.cat([s[:, 1] for s in shards]).long()
            assert ids.numel() == n - n % world and torch.unique(ids).numel() == ids.numel()      # disjoint
            assert all(torch.equal(s[:, 2], s[:, 1] * 3) and torch.equal(s[:, 0], s[:, 1] % 16) for s in shards)  # rows intact
            again = shard_sites(table, world - 1, world, seed=5)
            assert torch.equal(again, shards[-1])                                                  # deterministic
            if n >= 101 and world > 1:
                other = shard_sites(table, 0, world, seed=6)
                assert not torch.equal(other, shards[0])                                          # the seed matters
            # keeping the remainder: every site appears, shard lengths differ by at most one
            full = [shard_sites(table, r, world, seed=5, drop_remainder=False) for r in range(world)]
            assert sum(len(s) for s in full) == n and max(len(s) for s in full) - min(len(s) for s in full) <= 1
    # the shuffled shard is statistically like the whole table: per-image counts within 2 % on the C5-sized table
    table = torch.stack([torch.arange(2_000_000) % 16, torch.arange(2_000_000)], 1)
    counts = torch.bincount(shard_sites(table, 3, 8, seed=5)[:, 0], minlength=16).double()
    assert float((counts / counts.mean() - 1).abs().max()) < 0.02
