"""Seeded synthetic IMU windows and the window builder for the reference's ``.dat`` stream.

Host-side numpy only (input preparation is not on the hot path).  The generator follows SURVEY.md section 8(d):

* true motion per window: w(t) = w0 + A_w * sin(2 pi f_w t + phi_w), specific force a(t) = R_GtoI(t) ([0,0,9.8] + a_G(t)),
  R_GtoI integrated from w(t)  (at rest the accelerometer reads +9.8 in z, as in the reference datasets and
  solvers/GraphSolver.cpp:315);
* noise exactly as the reference simulator (cpi_simulation/SCRIPT_gazebo_to_sim.m:143-149):
  w_m = w + b + sigma/sqrt(dt) N(0,1),  b += sigma_w sqrt(dt) N(0,1);
  sigmas = (0.005, 4e-6, 0.01, 0.0002)  (Sigma.dat / launch/synthetic_test.launch:13-17);
* dt = 1/rate with 5 % of the steps doubled (the datasets drop samples: at 200 Hz 5 ms x10411, 10 ms x492, ...);
* 1 % of the windows forced into the ``small_w`` Taylor branch (|w_hat| in [1e-4, 0.008] rad/s, CpiV1.h:101),
  0.1 % with exactly zero w_hat, 0.1 % containing a dt = 0 step (CpiV1.h:72-74); every other sample is pushed out of
  the ill-conditioned band |w_hat| < 0.05 rad/s (SURVEY.md section 8a conditioning table).

Windows are generated in independent blocks of ``BLOCK`` windows, each from its own Philox stream keyed by
(seed, block index), so any rank can generate exactly its shard and the union is independent of the partition.
"""
from __future__ import annotations

import numpy as np

SEED = 20260924
SIGMAS = np.array([0.005, 4e-6, 0.01, 0.0002])   # sigma_w, sigma_wb, sigma_a, sigma_ab
GRAVITY = np.array([0.0, 0.0, 9.8])
SMALL_W = 0.008726646
BLOCK = 1024


def _rot_step(w, dt):
    """Rodrigues rotation exp(-[w dt]x) (JPL / global-to-local convention), batched: w (n,3) -> (n,3,3)."""
    th = np.linalg.norm(w, axis=1) * dt
    k = np.zeros((w.shape[0], 3, 3))
    k[:, 0, 1], k[:, 0, 2] = -w[:, 2], w[:, 1]
    k[:, 1, 0], k[:, 1, 2] = w[:, 2], -w[:, 0]
    k[:, 2, 0], k[:, 2, 1] = -w[:, 1], w[:, 0]
    n = np.maximum(np.linalg.norm(w, axis=1), 1e-300)
    a = np.where(th < 1e-8, dt, np.sin(th) / n)
    b = np.where(th < 1e-8, 0.5 * dt * dt, (1 - np.cos(th)) / (n * n))
    eye = np.eye(3)[None]
    return eye - a[:, None, None] * k + b[:, None, None] * (k @ k)


def _draw(seed, block, nb, entries):
    """All random draws of one block, in a fixed order, from the block's own Philox stream."""
    g = np.random.Generator(np.random.Philox(key=[seed, block]))
    d = {}
    d["w0"] = g.uniform(-1, 1, (nb, 3)); d["Aw"] = g.uniform(0, 1.5, (nb, 3)); d["fw"] = g.uniform(0.2, 2, (nb, 3)); d["pw"] = g.uniform(0, 2 * np.pi, (nb, 3))
    d["Aa"] = g.uniform(0, 2, (nb, 3)); d["fa"] = g.uniform(0.2, 2, (nb, 3)); d["pa"] = g.uniform(0, 2 * np.pi, (nb, 3))
    d["b_w"] = g.normal(0, 1e-3, (nb, 3)); d["b_a"] = g.normal(0, 1e-2, (nb, 3))
    q = g.normal(0, 1, (nb, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True); q[q[:, 3] < 0] *= -1
    d["q"] = q
    d["dsel"] = g.uniform(size=(nb, entries))
    d["nz_w"] = g.normal(size=(nb, entries, 3)); d["nz_a"] = g.normal(size=(nb, entries, 3))
    d["rw_w"] = g.normal(size=(nb, entries, 3)); d["rw_a"] = g.normal(size=(nb, entries, 3))
    d["kind"] = g.uniform(size=nb)
    d["scale"] = g.uniform(1e-4, 0.008, size=(nb, entries))
    d["zpos"] = g.integers(0, max(entries, 1), size=nb)
    d["bw_off"] = g.normal(0, 1e-4, (nb, 3)); d["ba_off"] = g.normal(0, 1e-3, (nb, 3))
    return d


def _blocks(seed, blocks, nb, ns, rate, special, entries):
    """Windows of several blocks at once: the draws come from each block's own stream, the time loop runs over all of them
    together (every operation is per window, so the result does not depend on how blocks are grouped)."""
    ds = [_draw(seed, b, nb, entries) for b in blocks]
    D = {k: np.concatenate([d[k] for d in ds]) for k in ds[0]}
    n = nb * len(blocks)
    dt0 = 1.0 / rate
    w0, Aw, fw, pw, Aa, fa, pa, b_w, b_a = (D[k] for k in ("w0", "Aw", "fw", "pw", "Aa", "fa", "pa", "b_w", "b_a"))
    lin = np.concatenate([b_w, b_a, D["q"], np.broadcast_to(GRAVITY, (n, 3))], axis=1)
    dts = np.where(D["dsel"] < 0.05, 2 * dt0, dt0)
    nz_w, nz_a, rw_w, rw_a, kind, scale, zpos = (D[k] for k in ("nz_w", "nz_a", "rw_w", "rw_a", "kind", "scale", "zpos"))

    S = np.zeros((n, entries, 7))
    R = np.broadcast_to(np.eye(3), (n, 3, 3)).copy()
    t = np.zeros(n)
    bw = b_w + D["bw_off"]                     # true bias near (not at) the linearisation point
    ba = b_a + D["ba_off"]
    for i in range(entries):
        dt = dts[:, i]
        w = w0 + Aw * np.sin(2 * np.pi * fw * t[:, None] + pw)
        aG = Aa * np.sin(2 * np.pi * fa * t[:, None] + pa)
        a = np.einsum("nij,nj->ni", R, GRAVITY[None] + aG)
        wm = w + bw + (SIGMAS[0] / np.sqrt(dt))[:, None] * nz_w[:, i]
        am = a + ba + (SIGMAS[2] / np.sqrt(dt))[:, None] * nz_a[:, i]
        bw = bw + SIGMAS[1] * np.sqrt(dt)[:, None] * rw_w[:, i]
        ba = ba + SIGMAS[3] * np.sqrt(dt)[:, None] * rw_a[:, i]
        S[:, i, 0:3], S[:, i, 3:6], S[:, i, 6] = wm, am, dt
        R = _rot_step(w, dt[:, None][:, 0]) @ R
        t = t + dt
    # keep ordinary samples out of the ill-conditioned band |w_hat| in [small_w, 0.05)
    what = S[:, :, 0:3] - b_w[:, None, :]
    mag = np.linalg.norm(what, axis=2)
    low = mag < 0.05
    if low.any():
        push = np.where(mag > 0, 0.05 * 1.5 / np.maximum(mag, 1e-300), 0.0)
        fix = what * push[:, :, None]
        fix[mag == 0] = np.array([0.075, 0.0, 0.0])
        S[:, :, 0:3] = np.where(low[:, :, None], fix + b_w[:, None, :], S[:, :, 0:3])
    if special:
        what = S[:, :, 0:3] - b_w[:, None, :]
        mag = np.linalg.norm(what, axis=2)
        small = kind < 0.01                                  # whole window in the Taylor branch
        S[small, :, 0:3] = (what[small] * (scale[small] / mag[small])[:, :, None]) + b_w[small, None, :]
        zero = (kind >= 0.01) & (kind < 0.011)               # exact-zero w_hat
        S[zero, :, 0:3] = b_w[zero, None, :]
        dz = (kind >= 0.011) & (kind < 0.012)                # one dt = 0 step
        S[dz, zpos[dz], 6] = 0.0
    return S, lin


def usable_cpus():
    """Host threads this process may actually use: scheduler affinity capped by the cgroup CPU quota (v2 cpu.max or v1 cfs)."""
    import math
    import os
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, math.ceil(int(quota) / int(period))))
    except (OSError, ValueError):
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read()); per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, math.ceil(q / per)))
        except (OSError, ValueError):
            pass
    return max(1, n)


def make_windows(n_windows, ns, rate=200.0, seed=SEED, first_window=0, special=True, imu_avg=False):
    """Return (samples[n, entries, 7], lin[n, 13]) for windows first_window .. first_window+n_windows-1.

    entries = ns (+1 trailing entry if imu_avg).  Deterministic in (seed, absolute window index, ns, rate).  Blocks are
    independent Philox streams; groups of 4 blocks share one pass of the per-sample time loop, groups run on a thread pool.
    """
    entries = ns + (1 if imu_avg else 0)
    b0, b1 = first_window // BLOCK, (first_window + n_windows + BLOCK - 1) // BLOCK
    Ss, Ls = [], []
    groups = [list(range(g0, min(g0 + 4, b1))) for g0 in range(b0, b1, 4)]
    gen = lambda grp: _blocks(seed, grp, BLOCK, ns, rate, special, entries)
    nthr = min(usable_cpus(), 16, len(groups))
    if nthr > 1:                                  # numpy releases the GIL on these array sizes
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(nthr) as ex:
            outs = list(ex.map(gen, groups))
    else:
        outs = [gen(g) for g in groups]
    blocks = []
    for grp, (Sg, Lg) in zip(groups, outs):
        blocks += [(Sg[i * BLOCK:(i + 1) * BLOCK], Lg[i * BLOCK:(i + 1) * BLOCK]) for i in range(len(grp))]
    for b, (S, L) in zip(range(b0, b1), blocks):
        lo = max(first_window - b * BLOCK, 0)
        hi = min(first_window + n_windows - b * BLOCK, BLOCK)
        Ss.append(S[lo:hi]); Ls.append(L[lo:hi])
    if not Ss:
        return np.zeros((0, entries, 7)), np.zeros((0, 13))
    return np.ascontiguousarray(np.concatenate(Ss)), np.ascontiguousarray(np.concatenate(Ls))


# --------------------------------------------------------------------------------------------------------------
# window builder for the reference's .dat stream ("next" row: SURVEY.md section 8f rank 3)
# --------------------------------------------------------------------------------------------------------------

def parse_imu_dat(text_or_path):
    """Parse ``imu_data_*.dat`` lines "wx wy wz ax ay az <unused> t_ms" (sim/SimParser.h:148-175).
    Returns (t[n] seconds, w[n,3], a[n,3])."""
    if isinstance(text_or_path, str) and "\n" not in text_or_path:
        d = np.loadtxt(text_or_path)
    else:
        d = np.loadtxt(text_or_path.splitlines() if isinstance(text_or_path, str) else text_or_path)
    d = np.atleast_2d(d)
    return 1e-3 * d[:, 7], d[:, 0:3].copy(), d[:, 3:6].copy()


def cut_windows(t, w, a, update_times, imu_wait=0):
    """Replay the reference driver loop (solvers/GraphSolver_IMU.cpp:50-69) over an IMU stream: the C ABI's host-side window
    builder ``cpi_cut_windows`` (include/cpi_b200.h; cpi_b200/csrc/windows.cu).

    For every camera/update time (ascending), consume IMU readings while ``imu_times[1] <= updatetime`` -- one step
    feed_IMU(t0, t1, w0, a0) per reading with dt >= 0 -- then, if ``updatetime - imu_times[0] > 0``, one partial step
    feed_IMU(t0, updatetime, w0, a0) and set imu_times[0] = updatetime.  ``imu_wait`` > 0 adds the reference's initialisation
    (the first update that finds that many queued readings emits no window).  Returns (samples[total,7], offsets[n+1]) in the
    CSR layout of include/cpi_b200.h (imu_avg = False, as at the call site GraphSolver_IMU.cpp:45).
    """
    import ctypes
    from . import capi
    lib = capi.load()
    t = np.ascontiguousarray(t, dtype=np.float64); w = np.ascontiguousarray(w, dtype=np.float64).reshape(-1, 3)
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1, 3); ut = np.ascontiguousarray(update_times, dtype=np.float64)
    off = np.zeros(len(ut) + 1, dtype=np.int64)
    ne = ctypes.c_int64(0)
    P = lambda x: ctypes.c_void_p(x.ctypes.data)
    cap = len(t) + len(ut) + 1                     # every reading is fed at most once, plus one partial step per update
    S = np.zeros((cap, 7))
    nw = lib.cpi_cut_windows(len(t), P(t), P(w), P(a), len(ut), P(ut), int(imu_wait), cap, P(S), P(off), ctypes.byref(ne))
    if nw < 0:
        capi.check(int(nw))
    return S[:ne.value].copy(), off[:nw + 1].copy()


def make_states(records, lin, model, seed=SEED, perturb=True):
    """A chain of JPLNavStates x_0..x_n for n windows: integrate the records with the reference's prediction
    (solvers/GraphSolver_IMU.cpp:263-307) and perturb (theta 1e-2 rad, v 1e-1, p 1e-1, biases 1e-3) -- SURVEY 8d config 5.
    Pure numpy; used to make factor-evaluation test inputs."""
    g = np.random.Generator(np.random.Philox(key=[seed, 0xC0FFEE]))
    n = records.shape[0]
    X = np.zeros((n + 1, 16))
    X[0, 3] = 1.0
    X[0, 4:7] = lin[0, 0:3]; X[0, 10:13] = lin[0, 3:6]

    def q2R(q):
        x, y, z, w_ = q
        sk = np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]])
        v = np.array([x, y, z])
        return (2 * w_ * w_ - 1) * np.eye(3) - 2 * w_ * sk + 2 * np.outer(v, v)

    def qmul(q, p):
        sk = np.array([[0, -q[2], q[1]], [q[2], 0, -q[0]], [-q[1], q[0], 0]])
        Q = np.zeros((4, 4)); Q[:3, :3] = q[3] * np.eye(3) - sk; Q[:3, 3] = q[:3]; Q[3, :3] = -q[:3]; Q[3, 3] = q[3]
        r = Q @ p
        if r[3] < 0:
            r = -r
        return r / np.linalg.norm(r)

    for k in range(n):
        r = records[k]; dt = r[19]; grav = lin[k, 10:13]
        Rinv = q2R(np.array([-X[k, 0], -X[k, 1], -X[k, 2], X[k, 3]]))
        X[k + 1, 0:4] = qmul(r[0:4], X[k, 0:4])
        X[k + 1, 4:7] = X[k, 4:7]; X[k + 1, 10:13] = X[k, 10:13]
        if model == 1:
            X[k + 1, 7:10] = X[k, 7:10] - grav * dt + Rinv @ r[16:19]
            X[k + 1, 13:16] = X[k, 13:16] + X[k, 7:10] * dt - 0.5 * grav * dt * dt + Rinv @ r[13:16]
        else:
            X[k + 1, 7:10] = X[k, 7:10] + Rinv @ r[16:19]
            X[k + 1, 13:16] = X[k, 13:16] + X[k, 7:10] * dt + Rinv @ r[13:16]
    if perturb:
        dth = g.normal(0, 1e-2, (n + 1, 3))
        for k in range(n + 1):
            nrm = np.linalg.norm(dth[k])
            dq = np.concatenate([np.sin(nrm / 2) / nrm * dth[k], [np.cos(nrm / 2)]])
            X[k, 0:4] = qmul(dq, X[k, 0:4])
        X[:, 4:7] += g.normal(0, 1e-3, (n + 1, 3)); X[:, 10:13] += g.normal(0, 1e-3, (n + 1, 3))
        X[:, 7:10] += g.normal(0, 1e-1, (n + 1, 3)); X[:, 13:16] += g.normal(0, 1e-1, (n + 1, 3))
    return X
