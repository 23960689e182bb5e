"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding, the one-time weight
arena broadcast, the max-over-ranks timing reduction, and the reference arm of bench.py under a
multi-rank launch (rank 0 alone prints, the others exit 0 without work)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "csi-nn2_b200", "pyhost"))

import b200_dist  # noqa: E402


def test_shard_batch_partitions_the_batch():
    for batch in (0, 1, 7, 256, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            seen = np.zeros(batch, int)
            sizes = []
            for r in range(world):
                s, c = b200_dist.shard_batch(batch, world, r)
                seen[s:s + c] += 1
                sizes.append(c)
            assert (seen == 1).all(), (batch, world)
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        b200_dist.shard_batch(8, 2, 2)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # rank 0 holds the packed weights, the others an empty arena of the same size
        gen = torch.Generator().manual_seed(7)
        ref = torch.randint(0, 256, (4501504,), dtype=torch.uint8, generator=gen)  # MobileNetV1 arena size
        arena = ref.clone() if rank == 0 else torch.zeros_like(ref)
        b200_dist.broadcast_arena(arena, src=0)
        ok_bcast = bool(torch.equal(arena, ref))
        mx = b200_dist.max_over_ranks([1.0 + rank, 5.0 - rank])
        start, count = b200_dist.shard_batch(513, world, rank)
        q.put((rank, ok_bcast, mx, start, count))
    finally:
        dist.destroy_process_group()


def test_weight_broadcast_and_max_reduce_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), "arena differs after broadcast"
    assert all(r[2] == [2.0, 5.0] for r in res)
    assert [(r[3], r[4]) for r in res] == [(0, 257), (257, 256)]


def test_bench_reference_arm_multirank_only_rank0_prints():
    """`bench.py --impl reference` under a 2-rank launch: rank 1 exits 0 silently, rank 0 prints one
    JSON line carrying impl / cpu_baseline / e2e with zero copy bytes"""
    env = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "3"]
    r1 = subprocess.run(cmd, env=dict(env, RANK="1", LOCAL_RANK="1"), capture_output=True, text=True, timeout=300)
    assert r1.returncode == 0 and r1.stdout.strip() == ""
    r0 = subprocess.run(cmd, env=dict(env, RANK="0", LOCAL_RANK="0"), capture_output=True, text=True, timeout=600)
    assert r0.returncode == 0, r0.stderr[-2000:]
    line = json.loads(r0.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["metric"] == "MobileNetV1 int8 inferences/sec"
