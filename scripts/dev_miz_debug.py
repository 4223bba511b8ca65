"""dev: small MIZ cases on the GPU, each meant to be run under `timeout`."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import ebm_b200 as ebm
from helpers import oracle_miz, rel_err

mode, nt, dur, nmem = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
st = ebm.SpaceTime(180, nt, dur, "sin")
par = ebm.default_parameters("MIZ")
z = np.zeros(180)
init = lambda: ebm.Collection(Ei=z.copy(), Ew=z.copy(), h=z.copy(), D=z.copy(), phi=z.copy())
t0 = time.time()
if mode == "step":
    v = init()
    for k in range(nt):
        ebm.step("MIZ", st.t[k], 0.0, v, st, par)
    print("step ok", time.time() - t0, v["newton_iters"], v["E"][:3], flush=True)
else:
    r = ebm.integrate_ensemble("MIZ", st, [ebm.Forcing(0.0)] * nmem, [par] * nmem, [init() for _ in range(nmem)],
                               field_stride=1 if nmem == 1 else 0, strict=(mode == "strict"), lastonly=False)
    print(mode, "ok", time.time() - t0, "iters", r.newton_iters[:4], "nonconv", r.nonconv[:4], flush=True)
    o = oracle_miz(st, [ebm.Forcing(0.0)], [par], [init()], lastonly=False, raw=True)
    print("oracle iters", o["newton_iters"], flush=True)
    for k in ("Ei", "Ew", "h", "D", "phi", "T0"):
        print(k, "max err", rel_err(r.final[k][0], o[k][0]).max(), flush=True)
if mode != "step" and nmem == 1 and r.raw is not None:
    for k in range(min(6, nt)):
        for vi, v in enumerate(ebm.MIZ_VARS):
            e = rel_err(r.raw[0, k, vi], o["raw"][0, k, vi])
            if e.max() > 1e-13:
                j = int(e.argmax())
                print(f"step {k} {v}: err {e.max():.3e} at cell {j}: gpu {r.raw[0, k, vi, j]!r} ref {o['raw'][0, k, vi, j]!r}", flush=True)
