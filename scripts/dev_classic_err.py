"""Dev helper: where does the fast classic kernel differ most from the oracle?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import ebm_b200 as ebm
from helpers import oracle_classic, rel_err, warm_init, cold_init
from test_classic_gpu import _ensemble
nmem, nx, nt = 40, 180, 2000
st = ebm.SpaceTime(nx, nt, 3)
forcings, pars, inits = _ensemble(nmem, nx)
o = oracle_classic(st, forcings, pars, inits, lastonly=True, raw=True, seasonal=True)
r = ebm.integrate_ensemble("Classic", st, forcings, pars, inits, field_stride=3)
sel = np.arange(0, nmem, 3)
err = rel_err(r.raw, o["raw"][sel])
k, ti, v, j = np.unravel_index(err.argmax(), err.shape)
print("worst", err.max(), "member", sel[k], "step", ti, "var", v, "cell", j)
for dt_ in range(-3, 4):
    t = ti + dt_
    if 0 <= t < nt:
        print(t, "gpu E,T,h", r.raw[k, t, :, j], "ref", o["raw"][sel[k], t, :, j], "err", err[k, t, :, j])
print("count > 1e-9:", (err > 1e-9).sum(), "of", err.size, "; > 1e-10:", (err > 1e-10).sum(), "; >1e-11:", (err>1e-11).sum())
big = np.argwhere(err > 1e-10)
Eref = o["raw"][sel][:, :, 0, :]
print("min |E_ref| at cells with err>1e-10:", [float(np.abs(Eref[a, max(b-1,0):b+1, d]).min()) for a, b, c, d in big[:10]])
