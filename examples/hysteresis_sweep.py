#!/usr/bin/env python
"""Hysteresis loop of the classic EBM (the scientific product of BASELINE config C4), on one B200.

    python examples/hysteresis_sweep.py [--members 4096] [--years 200] [--devices 0,1,...]

Two branches of a forcing sweep F = -20..+20 W/m^2 -- a warm start (ice free) and a cold start (snowball) -- are
integrated to equilibrium as one ensemble; the last year's annual-mean hemispheric temperature and ice area of every
member are the x / y of the reference's `plot_seasonal` hysteresis diagram (src/plot.jl:173-190).  Everything
numerical runs in libebm_cuda.so (no CPU fallback); the arrays go through `integrate_arrays`, the array form of the
reference's `integrate` for large ensembles.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ebm_b200 as ebm  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, default=4096, help="members per branch")
    ap.add_argument("--years", type=int, default=200)
    ap.add_argument("--devices", default="", help="comma-separated GPU ids: several GPUs behind one call")
    a = ap.parse_args()

    n, nx = a.members, 100
    st = ebm.SpaceTime(nx, 2000, a.years)
    p = ebm.default_parameters("Classic")
    par = np.tile([p[k] for k in ebm.CLASSIC_PAR_ORDER], (2 * n, 1))           # every member: the default parameters
    F = np.concatenate([np.linspace(-20.0, 20.0, n)] * 2)                        # the same sweep on both branches
    forc = np.zeros((2 * n, 10))
    forc[:, :3] = F[:, None]                                                     # Forcing(F): base = peak = cool
    warm = np.arange(2 * n) < n
    state = {"E": np.where(warm[:, None], 98.0, -9.5) * np.ones((2 * n, nx)),    # warm start / cold start
             "Tg": np.where(warm[:, None], 10.0, -10.0) * np.ones((2 * n, nx))}
    devices = [int(d) for d in a.devices.split(",")] if a.devices else None

    t0 = time.time()
    res = ebm.integrate_arrays("Classic", st, forc, par, state, devices=devices)
    dt = time.time() - t0
    T_mean, ice_area = ebm.hysteresis_points(res.diag)                           # [nmem, dur]: annual means per year
    print(f"{2 * n} members x {a.years} years in {dt:.2f} s = {2 * n * a.years / dt:,.0f} member-years/s (host buffers, "
          f"{'1 GPU' if not devices else str(len(devices)) + ' GPUs'}); flagged members: {int((res.flags != 0).sum())}")
    print("   F      warm branch: T [C]  ice area     cold branch: T [C]  ice area")
    for k in range(0, n, max(1, n // 16)):
        print(f"{F[k]:6.1f}   {T_mean[k, -1]:12.2f} {ice_area[k, -1]:9.3f}   {T_mean[n + k, -1]:12.2f} {ice_area[n + k, -1]:9.3f}")
    bistable = np.abs(ice_area[:n, -1] - ice_area[n:, -1]) > 0.5
    if bistable.any():
        print(f"bistable range of the forcing: F = {F[:n][bistable].min():.2f} .. {F[:n][bistable].max():.2f} W/m^2")


if __name__ == "__main__":
    main()
