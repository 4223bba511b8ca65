"""dev: tiny runs of every kernel for compute-sanitizer (memcheck / racecheck)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import ebm_b200 as ebm

def classic(nmem, nx, nt, dur, stride, uniform=True, strict=False):
    st = ebm.SpaceTime(nx, nt, dur)
    p = ebm.default_parameters("Classic")
    pars = [p if uniform else ebm.Collection({**p, "D": 0.5 + 0.01 * (m % 7)}) for m in range(nmem)]
    forc = [ebm.Forcing(-5.0 + 10.0 * m / max(nmem - 1, 1)) for m in range(nmem)]
    init = [ebm.Collection(E=np.full(nx, 98.0 if m % 2 else -9.5), Tg=np.full(nx, 10.0 if m % 2 else -10.0)) for m in range(nmem)]
    r = ebm.integrate_ensemble("Classic", st, forc, pars, init, field_stride=stride, strict=strict)
    print("classic", nmem, nx, nt, dur, "uniform" if uniform else "general", "strict" if strict else "fast", float(np.nanmean(r.diag)), flush=True)

def miz(nmem, nx, nt, dur, stride, xfunc="sin", strict=False):
    st = ebm.SpaceTime(nx, nt, dur, xfunc)
    p = ebm.default_parameters("MIZ")
    z = np.zeros(nx)
    init = [ebm.Collection(Ei=z, Ew=z, h=z, D=z, phi=z) for _ in range(nmem)]
    r = ebm.integrate_ensemble("MIZ", st, [ebm.Forcing(0.0)] * nmem, [p] * nmem, init, field_stride=stride, strict=strict)
    print("miz", nmem, nx, nt, dur, xfunc, "strict" if strict else "fast", float(np.nanmean(r.diag)), int(r.newton_iters.sum()), flush=True)

classic(70, 100, 100, 2, 3)              # uniform kernel, ragged member count, field output
classic(37, 100, 100, 1, 0, uniform=False)   # general (bands) kernel
classic(9, 180, 100, 1, 2)               # wide grid -> bands kernel <16,16>
classic(3, 60, 100, 1, 1, strict=True)   # literal kernel
miz(9, 180, 100, 2, 4)                   # fast kernel K=6
miz(5, 100, 100, 1, 1, "identity")       # K=4, identity grid
miz(3, 250, 100, 1, 0)                   # K=8
miz(2, 50, 100, 1, 1, strict=True)       # literal kernel
st = ebm.SpaceTime(180, 2000, 1, "sin"); v = ebm.Collection(Ei=np.zeros(180), Ew=np.zeros(180), h=np.zeros(180), D=np.zeros(180), phi=np.zeros(180))
ebm.step("MIZ", st.t[0], 0.0, v, st, ebm.default_parameters("MIZ")); print("miz step ok", flush=True)
print("fp64 peak", ebm.fp64_peak())
