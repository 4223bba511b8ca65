"""dev: host-API overhead of ebm_classic_run (H2D, reorder, transposes, D2H) at C4 size, 1 simulated year."""
import ctypes as C, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import ebm_b200 as ebm, bench
from ebm_b200 import _lib
lib = _lib.load()
nmem, years = 65536, int(sys.argv[1]) if len(sys.argv) > 1 else 1
bench.ORDER = sys.argv[2] if len(sys.argv) > 2 else "interleaved"
st, par, forc, (E0, Tg0) = bench.classic_workload(ebm, nmem, np.arange(nmem), years)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
h = [pin(par), pin(forc), pin(E0), pin(Tg0)]
diag = torch.empty((nmem, years, 3, 4), dtype=torch.float64).pin_memory()
Ef = torch.empty((nmem, 100), dtype=torch.float64).pin_memory(); Tgf = torch.empty((nmem, 100), dtype=torch.float64).pin_memory()
P = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_double))
out = _lib.ClassicOutputs(P(diag), None, None, P(Ef), P(Tgf), None)
grid = _lib.make_grid(st); opt = _lib.make_options()
for k in range(4):
    t0 = time.perf_counter()
    _lib.check(lib.ebm_classic_run(C.byref(grid), nmem, P(h[0]), P(h[1]), P(h[2]), P(h[3]), C.byref(opt), C.byref(out)))
    print(f"call {k}: {time.perf_counter() - t0:.3f} s", flush=True)
