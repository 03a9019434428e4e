"""Multi-GPU plumbing: image-tile partition + request-buffer exchange (not in the reference).

One process per GPU (torch.distributed, NCCL over NVLink). Rays and pixels are independent, every rank holds a full
replica of the brick store, so the data path needs no collective. The only shared mutable state is the brick
residency protocol: all replicas must stream the same bricks in the same order so that slot numbers
(Scene.cpp:224-225) stay identical. Per frame (or per batch of frames when nothing streams) the ranks all-gather
their fixed-size request blocks {count, positions[queue_size][3]} (12.3 KB at the reference's queue size) and apply
the same deterministic merge.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import check


def merge_request_blocks(blocks, queue_size):
    """Specification of the merge (pure numpy; the CUDA kernel requests_merge_kernel implements the same):
    blocks: int32 [world, 1 + 3*queue_size] = {count, positions}. Ranks in order, entries in queue order, duplicates
    dropped (first occurrence wins). Returns (total_unique, kept_positions[min(total, queue_size)][3], dropped[...][3])."""
    blocks = np.asarray(blocks, dtype=np.int32).reshape(-1, 1 + 3 * queue_size)
    seen, kept, dropped = set(), [], []
    for r in range(blocks.shape[0]):
        n = min(int(np.uint32(blocks[r, 0])), queue_size)
        pos = blocks[r, 1:1 + 3 * n].reshape(n, 3)
        for p in pos:
            key = (int(p[0]), int(p[1]), int(p[2]))
            if key in seen:
                continue
            seen.add(key)
            (kept if len(kept) < queue_size else dropped).append(key)
    return len(seen), np.array(kept, np.int32).reshape(-1, 3), np.array(dropped, np.int32).reshape(-1, 3)


class RequestExchange:
    """All-gather + merge of the per-rank brick request blocks."""

    def __init__(self, queue_size, device, world_size=None, group=None):
        self.q = int(queue_size)
        self.device = torch.device(device)
        self.group = group
        self.world = world_size if world_size is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.block = torch.zeros(1 + 3 * self.q, dtype=torch.int32, device=self.device)
        self.gathered = torch.zeros(self.world, 1 + 3 * self.q, dtype=torch.int32, device=self.device)

    def all_gather_blocks(self, block):
        """block: int32 [1 + 3q] on self.device (CPU tensors work with the gloo backend). Returns [world, 1 + 3q]."""
        if self.world == 1:
            self.gathered[0].copy_(block)
        elif self.device.type == "cuda":
            dist.all_gather_into_tensor(self.gathered.view(-1), block, group=self.group)
        else:
            dist.all_gather(list(self.gathered.unbind(0)), block, group=self.group)
        return self.gathered

    def exchange(self, renderer):
        """Device path: pack this rank's block, all-gather over NCCL, merge on every rank (CUDA kernel)."""
        if self.world == 1:
            return  # a single replica: the local queue already is the merged queue
        lib = _lib.load()
        stream = torch.cuda.ExternalStream(renderer.stream, device=self.device)
        check(lib.bm_requests_pack(renderer.h, self.block.data_ptr()), "bm_requests_pack")
        with torch.cuda.stream(stream):
            self.all_gather_blocks(self.block)
        check(lib.bm_requests_merge(renderer.h, self.gathered.data_ptr(), self.world), "bm_requests_merge")
