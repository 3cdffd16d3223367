"""Dev probe: kernel time vs windows per SM (model 1, fp64, 200 samples), device-resident."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cpi_b200 import capi, synth
lib = capi.load()
ns = 200
S, L = synth.make_windows(148 * 96, ns)
dS, dL = torch.from_numpy(S).cuda(), torch.from_numpy(L).cuda()
out = torch.empty((148 * 96, 290), dtype=torch.float64, device="cuda")
sig = np.ascontiguousarray(synth.SIGMAS)
st = torch.cuda.current_stream()
for per_sm in (1, 32, 64, 68, 70, 96):
    n = 148 * per_sm
    def l():
        lib.cpi_preintegrate_batch(1, 64, n, None, ns, ctypes.c_void_p(dS.data_ptr()), ctypes.c_void_p(dL.data_ptr()), ctypes.c_void_p(sig.ctypes.data), 0, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(st.cuda_stream))
    for _ in range(2): l()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): l()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"  {per_sm:3d} windows/SM: {ms:.3f} ms -> {n/ms*1e3/1e6:.2f} M w/s")
