"""Multi-GPU sharding of a window batch.

Windows share nothing but the four sigmas, so the path shards embarrassingly: rank r of G preintegrates the contiguous
block ``[lo(r), hi(r))`` and the only exchange is ONE all-gather of the fixed-size result records, after which every
rank (in particular rank 0, where the solver lives) holds all records in window order.  The kernel writes its shard
straight into its slice of the gather buffer (in-place all-gather, no pack kernel).

On GPUs the whole step -- kernel into the gather slice, then the in-place NCCL all-gather on a communication stream --
is ONE call into the C ABI (``cpi_preintegrate_batch_sharded``, include/cpi_b200.h; `Communicator` below wraps it):
``torch.distributed`` is only used once, to hand rank 0's NCCL id to the other ranks.  The pure-Python path through
``torch.distributed`` collectives remains for the CPU tests (gloo, injected ``compute``).  Because the all-gather needs
equal slices, the batch is padded to ``G * per_rank`` windows; padding windows are zero-step windows (identity record)
that are dropped after the gather.
"""
from __future__ import annotations

import numpy as np


class Communicator:
    """One NCCL communicator per process, created through the C ABI (cpi_comm_create).  ``step()`` enqueues this rank's kernel
    and the in-place all-gather; the all-gather runs on the communicator's own stream, so the next step (into another gather
    buffer) overlaps it.  ``wait()`` orders the current torch stream behind the last all-gather."""

    def __init__(self, group=None):
        import ctypes
        import torch
        import torch.distributed as dist
        from . import capi
        self.lib = capi.load()
        self.capi = capi
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        ident = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            capi.check(self.lib.cpi_comm_unique_id(ctypes.c_void_p(ident.data_ptr())))
        if self.world > 1:
            obj = [ident.numpy().tobytes()]
            dist.broadcast_object_list(obj, src=0, group=group)
            ident = torch.frombuffer(bytearray(obj[0]), dtype=torch.uint8)
        self._id = ident
        h = ctypes.c_void_p()
        capi.check(self.lib.cpi_comm_create(ctypes.c_void_p(ident.data_ptr()), self.rank, self.world, ctypes.byref(h)))
        self.handle = h

    def step(self, model, samples, lin, sigmas, flags, gather, offsets=None, ns=None, stream=None):
        """samples / lin / offsets: this rank's shard (torch CUDA tensors, float64 or float32); gather: (world, n_local, rd) tensor."""
        import ctypes
        import torch
        n_local = lin.shape[0]
        dtype = 32 if samples.dtype == torch.float32 else 64
        assert gather.is_contiguous() and gather.shape[0] == self.world and gather.shape[1] == n_local and gather.dtype == samples.dtype
        sig = np.ascontiguousarray(sigmas, dtype=np.float64)
        st = stream if stream is not None else torch.cuda.current_stream()
        P = lambda t: ctypes.c_void_p(t.data_ptr() if t is not None and t.numel() else 0)
        self.capi.check(self.lib.cpi_preintegrate_batch_sharded(self.handle, model, dtype, n_local, P(offsets), 0 if ns is None else ns, P(samples), P(lin),
                                                                ctypes.c_void_p(sig.ctypes.data), flags, P(gather), ctypes.c_void_p(st.cuda_stream)))

    def register(self, gather):
        """Collective, once per gather buffer (same order on every rank): lets step() exchange the records with copy-engine peer copies
        over NVLink instead of an NCCL all-gather kernel (cpi_comm_register).  Returns whether that path is active."""
        import ctypes
        flag = ctypes.c_int(0)
        self.capi.check(self.lib.cpi_comm_register(self.handle, ctypes.c_void_p(gather.data_ptr()), gather.numel() * gather.element_size(), ctypes.byref(flag)))
        return bool(flag.value)

    def unregister(self, gather=None):
        """Drop the registration of one gather buffer (None: all).  Every rank must do this -- and the ranks must synchronise -- before
        any rank frees the buffer."""
        import ctypes
        self.capi.check(self.lib.cpi_comm_unregister(self.handle, ctypes.c_void_p(gather.data_ptr() if gather is not None else 0)))

    def wait(self, stream=None):
        import ctypes
        import torch
        st = stream if stream is not None else torch.cuda.current_stream()
        self.capi.check(self.lib.cpi_comm_wait(self.handle, ctypes.c_void_p(st.cuda_stream)))

    def close(self):
        if self.handle:
            self.lib.cpi_comm_destroy(self.handle)
            self.handle = None


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


def preintegrate_sharded(model, samples, lin, sigmas, flags=0, offsets=None, ns=None, group=None, compute=None, comm=None, dtype=np.float64):
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
    dtype = np.dtype(dtype)
    if dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
        raise ValueError("dtype must be float64 or float32")
    lin = np.ascontiguousarray(lin, dtype=dtype).reshape(-1, 13)
    samples = np.ascontiguousarray(samples, dtype=dtype).reshape(-1, 7)
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
            s_loc = np.concatenate([s_loc, np.zeros((avg * pad, 7), dtype=dtype)])
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
            s_loc = np.concatenate([s_loc, np.zeros((avg * pad, 7), dtype=dtype)])
            ns_loc = None
    l_loc = lin[lo:hi]
    if per - (hi - lo):
        l_loc = np.concatenate([l_loc, np.zeros((per - (hi - lo), 13), dtype=dtype)])

    use_cuda = compute is None
    tdt = torch.float32 if dtype == np.dtype(np.float32) else torch.float64
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    gather = torch.empty((world, per, rd), dtype=tdt, device=dev)
    mine = gather[rank]
    if use_cuda:
        # a rank whose shard holds no real window still takes part: zero-step windows, and a one-entry sample buffer so that no
        # NULL pointer reaches the C ABI (a local failure before the collective would leave the other ranks hanging in it)
        if s_loc.shape[0] == 0:
            s_loc = np.zeros((1, 7), dtype=dtype)
        d_s = torch.from_numpy(np.ascontiguousarray(s_loc)).to(dev)
        d_l = torch.from_numpy(np.ascontiguousarray(l_loc)).to(dev)
        d_o = torch.from_numpy(loc).to(dev) if loc is not None else None
        own = comm is None
        if own:
            comm = Communicator(group)
        comm.step(model, d_s, d_l, sigmas, flags, gather, offsets=d_o, ns=ns_loc)
        comm.wait()
        if own:
            torch.cuda.current_stream().synchronize()
            comm.close()
    else:
        # CPU stand-in path (gloo tests): a failing rank must not leave the others blocked in the collective
        err = None
        try:
            compute(model, s_loc, l_loc, sigmas, flags, loc, ns_loc, mine)
        except Exception as e:      # noqa: BLE001 -- re-raised below on every rank
            err = e
        if world > 1:
            flag = torch.tensor([1 if err is not None else 0], dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
            if int(flag.item()):
                raise RuntimeError(f"preintegrate_sharded: the local computation failed on some rank (this rank: {err!r})")
            dist.all_gather_into_tensor(gather.view(-1), mine.reshape(-1), group=group)
        elif err is not None:
            raise err
    return gather.view(world * per, rd)[:n]
