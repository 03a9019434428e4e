"""CPU-only, world_size 2 over gloo: the request-block exchange that keeps brick-store replicas in lock step.
Both ranks must end up with the identical merged request list, whatever each of them asked for locally."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from brickmap_b200.parallel import RequestExchange, merge_request_blocks

Q = 16


def local_block(rank):
    rng = np.random.default_rng(100 + rank)
    n = [11, 14][rank]
    pos = rng.integers(0, 6, size=(n, 3)).astype(np.int32)
    pos = pos[np.sort(np.unique(pos, axis=0, return_index=True)[1])]  # a rank never requests a cell twice (requested bit)
    block = np.zeros(1 + 3 * Q, np.int32)
    block[0] = pos.shape[0]
    block[1:1 + 3 * pos.shape[0]] = pos.reshape(-1)
    return block


def worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ex = RequestExchange(Q, "cpu", world)
    gathered = ex.all_gather_blocks(torch.from_numpy(local_block(rank))).numpy().copy()
    total, kept, dropped = merge_request_blocks(gathered, Q)
    out[rank] = (gathered, total, kept, dropped)
    dist.barrier()
    dist.destroy_process_group()


def test_request_exchange_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(worker, args=(2, port, out), nprocs=2, join=True)
    g0, t0, k0, d0 = out[0]
    g1, t1, k1, d1 = out[1]
    assert np.array_equal(g0, g1) and t0 == t1 and np.array_equal(k0, k1) and np.array_equal(d0, d1)
    assert np.array_equal(g0[0], local_block(0)) and np.array_equal(g0[1], local_block(1))
    # merge semantics: rank order, queue order, duplicates dropped, capped at the queue size
    b0, b1 = local_block(0), local_block(1)
    seq = [tuple(p) for p in b0[1:1 + 3 * b0[0]].reshape(-1, 3)] + [tuple(p) for p in b1[1:1 + 3 * b1[0]].reshape(-1, 3)]
    uniq = list(dict.fromkeys(seq))
    assert t0 == len(uniq) and len(uniq) > Q, "the case must overflow the queue to cover the drop path"
    assert [tuple(p) for p in k0] == uniq[:Q] and [tuple(p) for p in d0] == uniq[Q:]


def test_merge_single_rank_is_identity():
    b = local_block(0)
    total, kept, dropped = merge_request_blocks(b[None], Q)
    assert total == b[0] and dropped.shape[0] == 0
    assert np.array_equal(kept.reshape(-1), b[1:1 + 3 * b[0]])


def test_merge_clamps_overflowing_count():
    b = np.zeros((1, 1 + 3 * Q), np.int32)
    b[0, 0] = 5 * Q  # the device counter may exceed the queue size (voxel.cuh:234-240); consumers clamp (kernel.cu:409)
    b[0, 1:] = np.arange(3 * Q)
    total, kept, _ = merge_request_blocks(b, Q)
    assert total == Q and kept.shape == (Q, 3)
