"""dev: classic ensemble on a 180-cell grid (uniform parameters): throughput through the host API."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import ebm_b200 as ebm
nmem, years = int(sys.argv[1]), int(sys.argv[2])
st = ebm.SpaceTime(180, 2000, years)
p = ebm.default_parameters("Classic")
par = np.tile([p[k] for k in ebm.CLASSIC_PAR_ORDER], (nmem, 1))
if len(sys.argv) > 3: par[:, 0] = np.linspace(0.45, 0.75, nmem)
forc = np.zeros((nmem, 10)); F = np.linspace(-10, 10, nmem); forc[:, 0] = forc[:, 1] = forc[:, 2] = F
warm = (np.arange(nmem) % 2) == 0
state = {"E": np.where(warm[:, None], 98.0, -9.5) * np.ones((nmem, 180)), "Tg": np.where(warm[:, None], 10.0, -10.0) * np.ones((nmem, 180))}
for k in range(2):
    t0 = time.perf_counter(); r = ebm.integrate_arrays("Classic", st, forc, par, state); dt = time.perf_counter() - t0
print(f"nx=180 {'D sweep' if len(sys.argv) > 3 else 'uniform'}: {nmem} x {years} y: {nmem * years / dt:.0f} member-years/s, flags {int(r.flags.max())}, mean T {np.nanmean(r.diag[:, -1, 2, 0]):.5f}")
