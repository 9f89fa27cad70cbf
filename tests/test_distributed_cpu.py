"""world_size-2 gloo tests (CPU) of the multi-GPU protocol: chains sharded by global chain id,
per-rank MergeChains contributions and within/between sums all-reduced, finalised through the
C ABI's host entry points.  The device kernels are stood in for by the oracle's device-schedule
restatement (bit-identical to them on a GPU: tests/test_gpu_parity.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import grample_b200 as gb
import oracle
from conftest import RES
from grample_b200 import distributed as gbd

TOTAL_CHAINS, SEED, CW = 20, 11, 6


def test_shard_covers_all_chains_block_aligned():
    for total in (1, 7, 8, 20, 4096, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [gbd.shard(total, world, r) for r in range(world)]
            assert sum(n for _, n in spans) == total
            pos = 0
            for first, n in spans:
                assert first % gbd.CHAIN_BLOCK == 0
                if n:
                    assert first == pos
                    pos += n


def test_library_shards_like_the_python_plumbing():
    """gb_shard — the rule the library applies to a new variant's chains under a communicator or a fleet
    (gb_chains_adapt / gb_fleet_adapt) — is the rule `distributed.shard` applies to the base chains"""
    import ctypes as C
    from grample_b200 import _lib
    for total in (1, 7, 8, 20, 64, 4096, 65536, 65537):
        for world in (1, 2, 3, 8):
            for rank in range(world):
                first, n = C.c_uint64(), C.c_int32()
                _lib.check(_lib.lib().gb_shard(total, world, rank, C.byref(first), C.byref(n)))
                assert (first.value, n.value) == gbd.shard(total, world, rank)
    first, n = C.c_uint64(), C.c_int32()
    assert _lib.lib().gb_shard(8, 2, 2, C.byref(first), C.byref(n)) != 0


def simulate_rank(first, n):
    """what one rank's device would hold after one advance(CW) round over chains [first, first+n)"""
    path = os.path.join(RES, "Grids_11.uai")
    dm = gb.Model.from_uai(path, device=-1)
    order, _ = dm.schedule()
    om = oracle.Model.load(path)
    samp = oracle.Sampler(oracle.Generator(1), om)
    cards = dm.cards
    st = np.array([[oracle.philox_init_value(SEED, first + c, v, int(cards[v])) for v in range(dm.n_vars)]
                   for c in range(n)], dtype=np.int32).reshape(n, dm.n_vars)
    counts = np.zeros(int(cards.sum()))
    seqs = np.zeros((n, dm.n_vars, CW + 1), dtype=np.int32)
    for s in range(CW + 1):
        if n:
            st, counts = samp.sweep_run(order, SEED, first, st, s, 1, bits=53, record=True, counts=counts)
            seqs[:, :, s] = st
    prior = np.concatenate([np.full(c, n / c) for c in cards])
    return dm, counts + prior, seqs[:, :, 1:]  # the window is the last CW samples


def within_between(dm, seqs, merged, which):
    """per-variable sums over this rank's chains of ChainDist's (within, between)"""
    cards = dm.cards
    offs = np.concatenate([[0], np.cumsum(cards)])
    wb = np.zeros(2 * dm.n_vars)
    for c in range(seqs.shape[0]):
        ch = oracle.Chain.from_marginals(cards, [np.ones(k) for k in cards], cw=CW)
        for v in range(dm.n_vars):
            ch.set_history(v, seqs[c, v])
            w, b = ch.chain_dist(which, v, merged[offs[v]:offs[v + 1]])
            wb[v] += w
            wb[dm.n_vars + v] += b
    return wb


def worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, n = gbd.shard(TOTAL_CHAINS, world, rank)
    dm, partial, seqs = simulate_rank(first, n)
    merged = gbd.all_reduce_numpy(dist, partial)          # MergeChains across ranks
    wb = gbd.all_reduce_numpy(dist, within_between(dm, seqs, merged, gb.HELLINGER))
    conv = gb.core.convergence_finalize(dm, wb, CW, TOTAL_CHAINS, np.zeros(dm.n_vars, np.int32))
    if rank == 0:
        np.save(out + ".merged.npy", merged)
        np.save(out + ".conv.npy", conv)
    dist.destroy_process_group()


def test_two_ranks_equal_one_rank(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "r")
    mp.spawn(worker, args=(2, port, out), nprocs=2, join=True)
    merged2, conv2 = np.load(out + ".merged.npy"), np.load(out + ".conv.npy")
    # single rank over all chains
    dm, merged1, seqs = simulate_rank(0, TOTAL_CHAINS)
    assert np.array_equal(merged1, merged2)  # counts are integers and the prior n/card sums exactly
    wb = within_between(dm, seqs, merged1, gb.HELLINGER)
    conv1 = gb.core.convergence_finalize(dm, wb, CW, TOTAL_CHAINS, np.zeros(dm.n_vars, np.int32))
    assert np.allclose(conv1, conv2, rtol=1e-12)
    # and the finalised scores are the reference's ChainConvergence over all 20 chains
    cards = dm.cards
    offs = np.concatenate([[0], np.cumsum(cards)])
    chains = []
    for c in range(TOTAL_CHAINS):
        marg = [merged1[offs[v]:offs[v + 1]] if c == 0 else np.zeros(cards[v]) for v in range(dm.n_vars)]
        ch = oracle.Chain.from_marginals(cards, marg, cw=CW)
        for v in range(dm.n_vars):
            ch.set_history(v, seqs[c, v])
        chains.append(ch)
    ref = oracle.chain_convergence(chains, oracle.HELLINGER, dm.n_vars)
    assert np.allclose(conv2, ref, rtol=1e-9)
