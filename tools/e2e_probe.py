import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cpi_b200 import capi, synth
lib = capi.load()
n, ns = 10000, 200
S, L = synth.make_windows(n, ns)
hS = torch.from_numpy(S).pin_memory(); hL = torch.from_numpy(L).pin_memory(); hO = torch.empty((n, 290), dtype=torch.float64).pin_memory()
dS = torch.empty_like(hS, device="cuda"); dO = torch.empty((n, 290), dtype=torch.float64, device="cuda")
sig = np.ascontiguousarray(synth.SIGMAS)
def t(f, k=10):
    for _ in range(3): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3
print("H2D 112 MB pinned: %.3f ms" % t(lambda: dS.copy_(hS, non_blocking=True)))
print("D2H 23 MB pinned: %.3f ms" % t(lambda: hO.copy_(dO, non_blocking=True)))
def host():
    capi.check(lib.cpi_preintegrate_batch_host(1, 64, n, None, ns, ctypes.c_void_p(hS.data_ptr()), ctypes.c_void_p(hL.data_ptr()), ctypes.c_void_p(sig.ctypes.data), 0, ctypes.c_void_p(hO.data_ptr())))
print("host call chunks=%s: %.3f ms" % (os.environ.get("CPI_B200_HOST_CHUNKS", "auto"), t(host)))
