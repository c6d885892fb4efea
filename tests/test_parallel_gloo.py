"""world_size-2 CPU test (gloo) of the multi-GPU exchange: the reference partition + ONE all_gather of the packed
(count | boxes) buffers, merged in rank order (lib/test.py:324-344); overflowing the payload raises."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from smallhardface_b200.parallel import gather_detections, merge_gathered, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_result(idx, cap=16):
    """Deterministic stand-in for image idx's post-vote boxes."""
    rng = np.random.RandomState(1000 + idx)
    n = int(rng.randint(0, cap))
    d = np.zeros((cap, 5), np.float32)
    d[:n] = rng.rand(n, 5)
    return d, n


def _worker(rank, world, port, n_images, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    per = int(np.ceil(n_images / world))
    a, b = shard_range(n_images, world, rank)
    dets = torch.zeros((per, 16, 5))
    cnts = torch.zeros((per,), dtype=torch.int32)
    for j, idx in enumerate(range(a, b)):
        d, n = _fake_result(idx)
        dets[j] = torch.from_numpy(d)
        cnts[j] = n
    calls = []
    orig = dist.all_gather
    dist.all_gather = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    g = gather_detections(dets, cnts, world)
    dist.all_gather = orig
    assert len(calls) == 1 and tuple(g.shape) == (world, per, 17, 5)          # one collective, counts ride in row 0
    merged = merge_gathered(g, n_images, world)
    if rank == 0:
        np.savez(os.path.join(out_dir, "merged.npz"), *merged)
        small = gather_detections(dets, cnts, world, rows=4)                  # (collective: both ranks must take part)
        try:
            merge_gathered(small, n_images, world)
            ok = all(_fake_result(i)[1] <= 4 for i in range(n_images))
        except RuntimeError as e:
            ok = "gather_rows" in str(e)
        assert ok
    else:
        gather_detections(dets, cnts, world, rows=4)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [5, 8])
def test_two_rank_all_gather_reproduces_single_process_order(tmp_path, n_images):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_images, str(tmp_path)), nprocs=2, join=True)
    z = np.load(os.path.join(str(tmp_path), "merged.npz"))
    merged = [z["arr_%d" % i] for i in range(len(z.files))]
    assert len(merged) == n_images
    for idx, got in enumerate(merged):
        d, n = _fake_result(idx)
        assert np.array_equal(got, d[:n])


def _cli_worker(rank, world, port, n_images, out_dir):
    """run_test.gather_all: the CLI's exchange of variable-length per-image results (float64 rows from box voting)."""
    from smallhardface_b200.run_test import gather_all
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard_range(n_images, world, rank)
    local = []
    for idx in range(a, b):
        d, n = _fake_result(idx, cap=40)
        local.append(d[:n].astype(np.float64))
    merged = gather_all(local, n_images, world, torch.device("cpu"))
    assert len(merged) == n_images
    for idx, got in enumerate(merged):
        d, n = _fake_result(idx, cap=40)
        assert got.dtype == np.float64 and np.array_equal(got, d[:n].astype(np.float64)), idx
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [3, 7])
def test_cli_gather_all_two_ranks(tmp_path, n_images):
    mp.spawn(_cli_worker, args=(2, _free_port(), n_images, str(tmp_path)), nprocs=2, join=True)
