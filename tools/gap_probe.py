"""Probe: where does the inter-launch gap come from?  (dev tool)"""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cpi_b200 import capi, preint, synth
lib = capi.load()
n, ns = 10000, 200
S, L = synth.make_windows(n, ns)
dS, dL = torch.from_numpy(S).cuda(), torch.from_numpy(L).cuda()
out = torch.empty((n, 290), dtype=torch.float64, device="cuda")
sig = np.ascontiguousarray(synth.SIGMAS)
st = torch.cuda.current_stream()
def launch():
    lib.cpi_preintegrate_batch(1, 64, n, None, ns, ctypes.c_void_p(dS.data_ptr()), ctypes.c_void_p(dL.data_ptr()), ctypes.c_void_p(sig.ctypes.data), 0, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(st.cuda_stream))
for _ in range(3): launch()
torch.cuda.synchronize()
for K in (1, 5, 20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(K): launch()
    t_issue = time.perf_counter() - t0
    e1.record(); torch.cuda.synchronize()
    print(f"K={K}: gpu {e0.elapsed_time(e1)/K:.3f} ms/launch, cpu issue {t_issue*1e3/K:.3f} ms/launch")
# smaller batches: time vs n
for nn in (148, 148*8, 148*32, 148*64, 148*68, 10000, 148*96, 20000, 40000):
    nn = min(nn, n)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    def l2():
        lib.cpi_preintegrate_batch(1, 64, nn, None, ns, ctypes.c_void_p(dS.data_ptr()), ctypes.c_void_p(dL.data_ptr()), ctypes.c_void_p(sig.ctypes.data), 0, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(st.cuda_stream))
    l2(); torch.cuda.synchronize()
    e0.record(); l2(); e1.record(); torch.cuda.synchronize()
    print(f"n={nn}: {e0.elapsed_time(e1):.3f} ms  -> {nn/e0.elapsed_time(e1)*1e3:.0f} win/s")
