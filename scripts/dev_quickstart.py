import sys; sys.path.insert(0, "/root/repo")
import numpy as np, ebm_b200 as ebm
st   = ebm.SpaceTime(100, 2000, 30)
par  = ebm.default_parameters("Classic")
init = ebm.Collection(E=np.full(100, 98.0), Tg=np.full(100, 10.0))
sols = ebm.integrate("Classic", st, ebm.Forcing(0.0), par, init)
forcings = [ebm.Forcing(F) for F in np.linspace(-20, 20, 4096)]
res = ebm.integrate_ensemble("Classic", st, forcings, [par] * 4096, [init] * 4096, field_stride=1024)
T_mean, ice_area = ebm.hysteresis_points(res.diag)
ebm.save_state("/tmp/run.ebm", res.final, years_done=30)
print(sols, T_mean.shape, float(T_mean[2048, -1]), float(ice_area[0, -1]), res.seasonal.shape)
