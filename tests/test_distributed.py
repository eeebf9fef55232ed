"""World-size-2 gloo test of the multi-GPU host logic: keyframe sharding + result gather.
Each rank runs the CPU oracle on its shard (the GPU path is exercised by bench.py --gpus N
on the box); the gathered records must equal a single-process run."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_shard(p, cfg, lo, hi):
    import orc
    from sloam_b200 import abi, capi
    n = hi - lo
    pts, mask = capi.synth_generate_host(cfg, lo, n)
    res = np.zeros(n, abi.KF_RESULT)
    empty_map, empty_planes = np.zeros(0, abi.CYLINDER), np.zeros(0, abi.PLANE)
    for k in range(n):
        pose = np.array([capi.synth_pose(cfg, lo + k)[1]])
        res[k] = orc.run_keyframe(p, pts[k], mask[k], pose, True, empty_map, empty_planes).result
    return res


def _worker(rank, world, port, n_keyframes, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sloam_b200 import capi, configs, sharding
    p, cfg = configs.make(capi, "os1-64")
    lo, hi = sharding.shard_range(n_keyframes, rank, world)
    local = _run_shard(p, cfg, lo, hi)
    full = sharding.gather_records(local, n_keyframes, rank, world, dist)
    if rank == 0:
        np.save(out_path, full.view(np.uint8))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition_the_sequence():
    from sloam_b200 import sharding
    for n, w in [(10, 2), (7, 4), (1000, 8), (3, 8)]:
        spans = [sharding.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_gather_equals_single_process(tmp_path):
    from sloam_b200 import abi, capi, configs
    n_keyframes, world = 5, 2          # uneven split: 3 + 2
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(world, 29000 + os.getpid() % 2000, n_keyframes, out), nprocs=world, join=True)
    got = np.load(out).view(abi.KF_RESULT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    p, cfg = configs.make(capi, "os1-64")
    exp = _run_shard(p, cfg, 0, n_keyframes)
    assert got.tobytes() == exp.tobytes()
    assert np.all(got["n_landmarks"] > 0)
