"""Multi-GPU plumbing: one process per GPU, chains sharded by GLOBAL chain id (the Philox stream
is keyed by it, so a run does not depend on the number of GPUs), and the only exchange is the
sum all-reduce of two small buffers at the monitor interval (cmd/root.go:498-539):
the MergeChains counts (sum(card) + 2 uint64) and the per-variable within/between sums
(2*n_vars + 1 float64).  On GPUs both reductions run INSIDE the library over its own NCCL
communicator (gb_comm_*: `attach` below hands rank 0's 128-byte id to the other ranks through
torch.distributed, which is used for nothing else on that path); the host-array helpers remain for
the gloo CPU tests of the protocol.
"""
import numpy as np

from . import core

CHAIN_BLOCK = 8  # chains share Philox calls in blocks of 8: shards start on multiples of 8


def shard(total_chains, world, rank):
    """(first_chain_id, n_chains) of `rank`: contiguous, block-aligned, covering [0, total)."""
    blocks = (total_chains + CHAIN_BLOCK - 1) // CHAIN_BLOCK
    per = (blocks + world - 1) // world
    first = min(rank * per, blocks) * CHAIN_BLOCK
    last = min((rank + 1) * per, blocks) * CHAIN_BLOCK
    return first, max(0, min(last, total_chains) - first)


class DeviceBuffer:
    """__cuda_array_interface__ view of a float64 device buffer owned by the library."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def all_reduce_numpy(dist, arr):
    """sum all-reduce of a host float64 array (gloo / CPU tests)"""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64).copy())
    dist.all_reduce(t)
    return t.numpy()


def all_reduce_device(dist, ptr, n, device):
    """in-place sum all-reduce of a library-owned device buffer (NCCL)"""
    import torch
    t = torch.as_tensor(DeviceBuffer(ptr, n), device=f"cuda:{device}")
    dist.all_reduce(t)
    torch.cuda.synchronize(device)


def attach(chains, dist):
    """Give `chains` an in-library NCCL communicator over the ranks of `dist` (no-op for one rank).  Collective."""
    if dist is None or dist.get_world_size() == 1 or getattr(chains, "comm", None) is not None:
        return getattr(chains, "comm", None)
    box = [core.Comm.unique_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    comm = core.Comm.init_rank(box[0], dist.get_world_size(), dist.get_rank(), chains.device)
    chains.attach_comm(comm)
    return comm


def merged_marginals(chains, dist=None, out=None):
    """sampler.MergeChains over the chains of every rank; `out` = optional caller-owned
    (float64 [sum(card)], int32 [n_vars]) host buffers to fill instead of allocating"""
    if dist is None or dist.get_world_size() == 1 or getattr(chains, "comm", None) is not None:
        return chains.merged_marginals(out)  # collective inside the library when a communicator is attached
    ptr, n = chains.merge_partial_dev()
    all_reduce_device(dist, ptr, n, chains.device)
    return chains.merge_finalize(out)


def convergence(chains, measure, merged, collapsed, cw, dist=None):
    """sampler.ChainConvergence over the chains of every rank"""
    if dist is None or dist.get_world_size() == 1 or getattr(chains, "comm", None) is not None:
        return chains.convergence(measure, merged)
    import torch
    ptr, n = chains.convergence_partial_dev(measure, merged)
    all_reduce_device(dist, ptr, n, chains.device)
    wb = torch.as_tensor(DeviceBuffer(ptr, n), device=f"cuda:{chains.device}").cpu().numpy()
    total = torch.tensor([chains.n_chains], dtype=torch.int64, device=f"cuda:{chains.device}")
    dist.all_reduce(total)
    return core.convergence_finalize(chains.base, wb, cw, int(total.item()), collapsed)


def adapt(chains, base_model, new_chain_count, replicas, cw, next_variant_id, measure, dist=None, max_groups=128):
    """(*ConvergenceSampler).Adapt over the chains of every rank.  Each new variant gets `replicas` chains
    globally, sharded like the base chains; `next_variant_id` is the global id of the new variants' first
    chain (consecutive variants are `stride` apart).  Every rank computes the same scores from the
    all-reduced sums, so every rank collapses the same variables.  Returns (chosen variables, stride)."""
    stride = (replicas + CHAIN_BLOCK - 1) // CHAIN_BLOCK * CHAIN_BLOCK
    if dist is None or dist.get_world_size() == 1 or getattr(chains, "comm", None) is not None:
        # with a communicator attached the library shards the new variants' chains itself (collective call)
        return chains.adapt(base_model, new_chain_count, replicas, cw, first_chain_id=next_variant_id, measure=measure,
                            max_groups=max_groups), stride
    world, rank = dist.get_world_size(), dist.get_rank()
    first, n = shard(replicas, world, rank)
    if n == 0:
        raise core.GrampleError("fewer than %d replicas per variant: rank %d would hold no chains" % (CHAIN_BLOCK * world, rank))
    if chains.n_groups >= max_groups:  # adaptive.go:62-64: nothing to do, so no reductions either
        return [], stride
    merged, col = merged_marginals(chains, dist)
    import torch
    total = torch.tensor([chains.n_chains], dtype=torch.int64, device=f"cuda:{chains.device}")
    dist.all_reduce(total)
    scores = convergence(chains, measure, merged, col, cw, dist)
    chosen = chains.adapt_scores(base_model, new_chain_count, n, scores, int(total.item()), next_variant_id + first,
                                 id_stride=stride, max_groups=max_groups)
    return chosen, stride


def total_samples(chains, dist=None):
    if dist is None or dist.get_world_size() == 1:
        return chains.total_samples

    import torch
    t = torch.tensor([chains.total_samples], dtype=torch.int64, device=f"cuda:{chains.device}")
    dist.all_reduce(t)
    return int(t.item())
