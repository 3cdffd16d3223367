"""N > 1 host logic on CPU: world_size-2/3 gloo runs of the shard + all-gather path with the oracle standing in for
the kernel (tests may use the oracle as the checker AND, here, as the stand-in compute: no GPU on this box)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cpi_b200 import shard, synth


def test_partition_covers_everything():
    for n in (0, 1, 7, 8, 9, 10000, 10001):
        for world in (1, 2, 3, 8):
            spans = [shard.partition(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(hi - lo <= per for lo, hi, per in spans)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, model, flags, ragged, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import Oracle
    orc = Oracle()
    n, ns = 13, 12
    S, L = synth.make_windows(n, ns, imu_avg=bool(flags & 1))
    ent = S.shape[1]
    if ragged:
        lens = np.array([(3 * i) % (ent + 1) for i in range(n)])
        if flags & 1:
            lens = np.maximum(lens, 1)
        off = np.zeros(n + 1, dtype=np.int64); off[1:] = np.cumsum(lens)
        Sx = np.concatenate([S[i, :lens[i]] for i in range(n)])
        kw = dict(offsets=off)
    else:
        Sx, kw = S.reshape(-1, 7), dict(ns=ns)

    def compute(model, s, l, sig, fl, loc, ns_loc, out):
        out.copy_(torch.from_numpy(orc.preintegrate(model, s, l, sig, fl, offsets=loc, ns=ns_loc)))

    got = shard.preintegrate_sharded(model, Sx, L, synth.SIGMAS, flags, compute=compute, **kw).numpy()
    ref = orc.preintegrate(model, Sx, L, synth.SIGMAS, flags, **kw)
    q.put((rank, bool(np.array_equal(got, ref)), got.shape))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("model,flags,ragged", [(1, 0, False), (2, 0, True), (1, 1, True), (2, 3, False)])
def test_sharded_gather_matches_unsharded(oracle, world, model, flags, ragged):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, model, flags, ragged, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res          # every rank holds every record, bit-identical to the unsharded run


def _worker_edge(rank, world, port, mode, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import Oracle
    orc = Oracle()
    n, ns = (2, 6) if mode == "empty_rank" else (7, 6)        # n = 2 on 3 ranks: ceil(2/3) = 1 -> rank 2 holds no real window
    S, L = synth.make_windows(n, ns)

    def compute(model, s, l, sig, fl, loc, ns_loc, out):
        if mode == "fail" and rank == 1:
            raise ValueError("injected failure on rank 1")
        out.copy_(torch.from_numpy(orc.preintegrate(model, s, l, sig, fl, offsets=loc, ns=ns_loc)))

    try:
        got = shard.preintegrate_sharded(1, S.reshape(-1, 7), L, synth.SIGMAS, 0, ns=ns, compute=compute).numpy()
        ok = bool(np.array_equal(got, orc.preintegrate(1, S.reshape(-1, 7), L, synth.SIGMAS, 0, ns=ns)))
        q.put((rank, "ok" if ok else "mismatch"))
    except RuntimeError as e:
        q.put((rank, "raised"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["empty_rank", "fail"])
def test_sharded_edge_cases(oracle, mode):
    """(a) a rank whose shard holds no real window still takes part in the collective (zero-step padding windows);
    (b) a local failure on one rank is raised on EVERY rank instead of leaving the others blocked in the all-gather."""
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_edge, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert set(res.values()) == ({"ok"} if mode == "empty_rank" else {"raised"}), res
