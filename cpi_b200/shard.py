"""Multi-GPU sharding of a window batch.

Windows share nothing but the four sigmas, so the path shards embarrassingly: rank r of G preintegrates the contiguous
block ``[lo(r), hi(r))`` and the only exchange is ONE all-gather of the fixed-size result records, after which every
rank (in particular rank 0, where the solver lives) holds all records in window order.  The kernel writes its shard
straight into its slice of the gather buffer (in-place all-gather, no pack kernel).

The collective goes through ``torch.distributed`` (NCCL over NVLink on GPUs; gloo in the CPU tests).  Because
``all_gather_into_tensor`` needs equal slices, the batch is padded to ``G * per_rank`` windows; padding windows are
zero-step windows (identity record) that are dropped after the gather.
"""
from __future__ import annotations

import numpy as np


def partition(n_windows: int, world: int, rank: int):
    """Contiguous block partition: returns (lo, hi, per_rank) with per_rank = ceil(n / world)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    per = (n_windows + world - 1) // world
    lo = min(rank * per, n_windows)
    hi = min(lo + per, n_windows)
    return lo, hi, per


def shard_csr(offsets, lo, hi):
    """Slice a CSR window layout: returns (first_entry, last_entry, local_offsets[hi-lo+1])."""
    offsets = np.asarray(offsets, dtype=np.int64)
    first, last = int(offsets[lo]), int(offsets[hi])
    return first, last, offsets[lo:hi + 1] - first


def preintegrate_sharded(model, samples, lin, sigmas, flags=0, offsets=None, ns=None, group=None, compute=None):
    """Preintegrate this rank's shard and all-gather the records.

    ``samples`` / ``lin`` / ``offsets`` describe the WHOLE batch (host numpy, identical on every rank) -- each rank
    touches only its slice.  Returns a torch tensor (n_windows, record_doubles) holding every window's record, on every
    rank.  ``compute(model, samples, lin, sigmas, flags, offsets, ns, out)`` fills ``out`` (a slice of the gather
    buffer) for the local shard; the default runs the CUDA kernel on the current device.  (The CPU tests inject a
    stand-in here to exercise the partition / gather logic under gloo.)
    """
    import torch
    import torch.distributed as dist

    from .capi import FLAG_IMU_AVG, REC_DOUBLES

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lin = np.ascontiguousarray(lin, dtype=np.float64).reshape(-1, 13)
    samples = np.ascontiguousarray(samples, dtype=np.float64).reshape(-1, 7)
    n = lin.shape[0]
    lo, hi, per = partition(n, world, rank)
    avg = 1 if flags & FLAG_IMU_AVG else 0
    rd = REC_DOUBLES[model]
    if offsets is not None:
        first, last, loc = shard_csr(offsets, lo, hi)
        s_loc = samples[first:last]
        # pad with zero-step windows up to per_rank (imu_avg windows carry one trailing entry even when empty)
        pad = per - (hi - lo)
        if pad:
            loc = np.concatenate([loc, loc[-1] + avg * np.arange(1, pad + 1, dtype=np.int64)])
            s_loc = np.concatenate([s_loc, np.zeros((avg * pad, 7))])
        ns_loc = None
    else:
        if ns is None:
            ns = samples.shape[0] // max(n, 1) - avg
        ent = ns + avg
        s_loc = samples[lo * ent:hi * ent]
        loc = None
        ns_loc = ns
        pad = per - (hi - lo)
        if pad:
            # uniform layout cannot express an empty window: switch the shard to CSR with zero-step padding
            loc = np.concatenate([np.arange(hi - lo + 1, dtype=np.int64) * ent, (hi - lo) * ent + avg * np.arange(1, pad + 1, dtype=np.int64)])
            s_loc = np.concatenate([s_loc, np.zeros((avg * pad, 7))])
            ns_loc = None
    l_loc = lin[lo:hi]
    if per - (hi - lo):
        l_loc = np.concatenate([l_loc, np.zeros((per - (hi - lo), 13))])

    use_cuda = compute is None
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    gather = torch.empty((world, per, rd), dtype=torch.float64, device=dev)
    mine = gather[rank]
    if use_cuda:
        from . import preint
        d_s = torch.from_numpy(np.ascontiguousarray(s_loc)).to(dev)
        d_l = torch.from_numpy(np.ascontiguousarray(l_loc)).to(dev)
        d_o = torch.from_numpy(loc).to(dev) if loc is not None else None
        preint.preintegrate(model, d_s, d_l, sigmas, flags, offsets=d_o, ns=ns_loc, out=mine)
    else:
        compute(model, s_loc, l_loc, sigmas, flags, loc, ns_loc, mine)
    if world > 1:
        dist.all_gather_into_tensor(gather.view(-1), mine.reshape(-1), group=group)
    return gather.view(world * per, rd)[:n]
