"""dev: find members whose final state is flagged non-finite in the C5-style sweep and re-run them on the oracle."""
import sys, ctypes as C
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import ebm_b200 as ebm, oracle, bench
from ebm_b200 import _lib

nmem, years = int(sys.argv[1]), int(sys.argv[2])
strict = len(sys.argv) > 3 and sys.argv[3] == "strict"
st, par, forc, init = bench.miz_workload(ebm, nmem, np.arange(nmem), years)
lib = _lib.load()
nx = st.nx
fin = [np.empty((nmem, nx)) for _ in range(6)]
flags = np.zeros(nmem, dtype=np.int32); it = np.zeros(nmem, dtype=np.int64); nc = np.zeros(nmem, dtype=np.int64)
out = _lib.MizOutputs(None, None, None, *[_lib.dptr(a) for a in fin], it.ctypes.data_as(C.POINTER(C.c_int64)),
                      nc.ctypes.data_as(C.POINTER(C.c_int64)), flags.ctypes.data_as(C.POINTER(C.c_int32)))
grid = _lib.make_grid(st); opt = _lib.make_options(strict=strict)
par = np.ascontiguousarray(par); forc = np.ascontiguousarray(forc)
_lib.check(lib.ebm_miz_run(C.byref(grid), nmem, _lib.dptr(par), _lib.dptr(forc), *[_lib.dptr(np.ascontiguousarray(a)) for a in init],
                           None, C.byref(opt), C.byref(out)))
bad = np.nonzero(flags)[0]
print("flagged", len(bad), bad[:10], "nonconv members", int((nc > 0).sum()), "max nonconv", int(nc.max()), "mean iters/step", it.mean() / (2000 * years))
for m in bad[:3]:
    print("member", m, "par D,B,ai,k,m1", par[m][[0, 2, 9, 11, 14]], "nonconv", nc[m])
    for k, a in zip(("Ei", "Ew", "h", "D", "phi", "T0"), fin):
        w = np.nonzero(~np.isfinite(a[m]))[0]
        print("  ", k, "nonfinite cells", w[:8], a[m][w[:4]])
    o = oracle.miz_run(st.x, st.t, years, st.winter.inx, st.summer.inx, 1, par[m:m+1], forc[m:m+1], *[a[m:m+1] for a in init])
    print("  oracle finite:", {k: bool(np.isfinite(o[k]).all()) for k in ("Ei", "Ew", "h", "D", "phi")}, "nonconv", o["nonconv"])
