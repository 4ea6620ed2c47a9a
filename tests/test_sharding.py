"""CPU: batch-sharding host logic with a real 2-process gloo group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ccvpe_b200.sharding import gather_poses, shard_bounds


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(n_total, rank, world)
    idx = torch.arange(lo, hi, dtype=torch.int64) * 10
    pose = dict(idx=idx, rc=torch.stack([idx // 512, idx % 512], dim=1).to(torch.int32),
                cs=torch.stack([idx.float().cos(), idx.float().sin()], dim=1),
                angle=idx.double() * 0.5, valid=(idx % 3 != 0).to(torch.uint8))
    full = gather_poses(pose, n_total)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **{k: v.numpy() for k, v in full.items()})
    dist.destroy_process_group()


def test_gather_poses_two_ranks_ragged(tmp_path):
    n_total, world = 7, 2                      # ragged: shards of 4 and 3 pairs
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    ref_idx = np.arange(n_total, dtype=np.int64) * 10
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        assert got["idx"].tolist() == ref_idx.tolist()                       # batch order restored on every rank
        assert got["rc"].tolist() == np.stack([ref_idx // 512, ref_idx % 512], 1).tolist()
        np.testing.assert_allclose(got["angle"], ref_idx * 0.5)
        assert got["valid"].tolist() == (ref_idx % 3 != 0).astype(np.uint8).tolist()
