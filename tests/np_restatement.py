"""Independent NumPy restatement of the reference's step functions (tests only).

Purpose: a second, structurally different reading of ``src/classic.jl`` and ``src/miz.jl``
(one Python function per Julia function, dense matrices, dense Newton Jacobian) used to
cross-check the C oracle in ``oracle/`` so that a shared misreading is less likely.  It is
NOT part of the product and is never imported outside ``tests/``.

Citations are relative to the reference repository.
"""
from __future__ import annotations

import numpy as np

PI = float(np.pi)


# --------------------------------------------------------------------------- shared pieces
def get_diffop(nx: int) -> np.ndarray:
    """src/infrastructure.jl:480-489 as a dense matrix."""
    dx = 1.0 / nx
    xb = np.array([j / nx for j in range(1, nx)])
    lam = (1 - xb**2) / dx**2
    l1 = np.concatenate([[0.0], -lam])
    l2 = np.concatenate([-lam, [0.0]])
    l3 = -l1 - l2
    return np.diag(-l1[1:nx], -1) + np.diag(-l3, 0) + np.diag(-l2[0:nx - 1], 1)


def generic_stencil_cache(x: np.ndarray):
    """src/infrastructure.jl:509-519."""
    nx = len(x)
    xe = np.concatenate([[-x[0]], x, [2 - x[-1]]])
    diffx = np.diff(xe)
    i = np.arange(1, nx + 1)
    xxph = (xe[i + 1] + xe[i]) / 2.0
    xxmh = (xe[i] + xe[i - 1]) / 2.0
    return dict(diffx=diffx, mxxph=1.0 - xxph**2, mxxmh=1.0 - xxmh**2, phmmh=xxph - xxmh, i=i)


def diffusion(temp: np.ndarray, x: np.ndarray, D: float, kind: int, cache=None) -> np.ndarray:
    """diffusion(T, st, par): identity grid (:495-497) or generic stencil (:505-526)."""
    nx = len(temp)
    if kind == 0:
        return (D * get_diffop(nx)) @ temp
    c = cache if cache is not None else generic_stencil_cache(x)
    diffT = np.zeros(nx + 1)
    diffT[1:nx] = np.diff(temp)
    i = c["i"]
    with np.errstate(all="ignore"):
        return D * (c["mxxph"] * diffT[i] / c["diffx"][i] - c["mxxmh"] * diffT[i - 1] / c["diffx"][i - 1]) / c["phmmh"]


def jl_min(a, b):
    """Julia min propagates NaN; np.minimum does too."""
    return np.minimum(a, b)


def crossmean(rows: np.ndarray) -> np.ndarray:
    """src/utilities.jl:390-395: per-cell mean over the year's nt vectors (rows [nt, nx])."""
    return rows.sum(axis=0) / rows.shape[0]


# --------------------------------------------------------------------------- classic
def classic_statics(x, t, nt, par, stencil=0):
    """get_statics (classic.jl:16-33).  ``stencil=1``: the generic flux-form stencil (infrastructure.jl:505-526) in kappa
    instead of get_diffop(nx) -- the extension for non-uniform grids."""
    nx = len(x)
    dt = 1.0 / nt
    cg_tau = par["cg"] / par["tau"]
    dt_tau = dt / par["tau"]
    dc = dt_tau * cg_tau
    dop = diffusion_matrix(x, 1.0, 1, generic_stencil_cache(x)) if stencil else get_diffop(nx)
    kappa = (1 + dt_tau) * np.eye(nx) - dt * par["D"] * dop / par["cg"]
    S = (par["S0"] - par["S2"] * x**2)[:, None] - (par["S1"] * np.cos(2.0 * PI * t))[None, :] * x[:, None]
    S = np.hstack([S, S[:, :1]])
    M = par["B"] + cg_tau
    aw = par["a0"] - par["a2"] * x**2
    kLf = par["k"] * par["Lf"]
    return dict(dt=dt, cg_tau=cg_tau, dt_tau=dt_tau, dc=dc, kappa=kappa, S=S, M=M, aw=aw, kLf=kLf)


def classic_step(stat, par, i, f, E, Tg, debug=None):
    """src/classic.jl:43-65; ``i`` is the 1-based year index.  Returns (E, Tg, T, h), plus the local named by ``debug``
    (the reference's `debug::Expr` is evaluated in this scope, :67-69)."""
    S = stat["S"]
    with np.errstate(all="ignore"):
        alpha = np.where(E > 0.0, stat["aw"], 0.0) + np.where(E < 0.0, par["ai"], 0.0)
        C = alpha * S[:, i - 1] + stat["cg_tau"] * Tg - par["A"] + f
        T0 = C / (stat["M"] - stat["kLf"] / E)
        T = np.where(E >= 0, E / par["cw"], 0.0) + np.where((E < 0.0) & (T0 < 0.0), T0, 0.0)
        E = E + stat["dt"] * (C - stat["M"] * T + par["Fb"])
        mask = (T0 < 0.0) & (E < 0.0)
        g = stat["M"] - stat["kLf"] / E
        A = stat["kappa"] - np.diag(np.where(mask, stat["dc"] / g, 0.0))
        rhs = Tg + stat["dt_tau"] * (np.where(E >= 0, E / par["cw"], 0.0)
                                     + np.where(mask, (par["ai"] * S[:, i] - par["A"] + f) / g, 0.0))
        Tg = np.linalg.solve(A, rhs)
        h = np.where(E < 0.0, -E / par["Lf"], 0.0)
    if debug is not None:
        return E, Tg, T, h, dict(alpha=alpha, C=C, T0=T0, S=S[:, i - 1], mask=mask.astype(float))[debug]
    return E, Tg, T, h


def classic_integrate(st, forcing, par, E0, Tg0, stencil=0):
    """integrate(:Classic, ...) storing every step (lastonly=false).  Returns dict of [nt*dur, nx]."""
    stat = classic_statics(st.x, st.t, st.nt, par, stencil)
    E, Tg = E0.copy(), Tg0.copy()
    n = st.nt * st.dur
    out = {k: np.empty((n, st.nx)) for k in ("E", "T", "h", "Tg")}
    for tinx in range(1, n + 1):
        ti = (tinx - 1) % st.nt + 1
        E, Tg, T, h = classic_step(stat, par, ti, forcing(st.T(tinx)), E, Tg)
        out["E"][tinx - 1], out["T"][tinx - 1], out["h"][tinx - 1], out["Tg"][tinx - 1] = E, T, h, Tg
    return out


# --------------------------------------------------------------------------- MIZ
def solar(x, t, ice, par):
    base = par["S0"] - par["S1"] * x * np.cos(2.0 * PI * t) - par["S2"] * x**2
    return par["ai"] * base if ice else (par["a0"] - par["a2"] * x**2) * base


def Tbar(Ti, Tw, phi):
    return Ti * phi + (1 - phi) * Tw


def T0eq(T0, x, t, hp, Tw, phi, f, par, kind, cache):
    vec = par["k"] * (par["Tm"] - T0) / hp
    vec = vec + solar(x, t, True, par)
    vec = vec + ((-par["A"]) - par["B"] * (T0 - par["Tm"]))
    vec = vec + diffusion(Tbar(jl_min(T0, par["Tm"]), Tw, phi), x, par["D"], kind, cache)
    return vec + f


def diffusion_matrix(x, D, kind, cache):
    nx = len(x)
    eye = np.eye(nx)
    return np.stack([diffusion(eye[:, k], x, D, kind, cache) for k in range(nx)], axis=1)


def solveTi(T0, x, t, h, Tw, phi, f, par, kind, cache, Lmat, tol=1e-8):
    """src/miz.jl:47-68 with a dense semi-smooth Newton in place of NonlinearSolve.TrustRegion."""
    hp = np.where(h == 0.0, par["hmin"], h)
    iters = 0
    while True:
        res = T0eq(T0, x, t, hp, Tw, phi, f, par, kind, cache)
        if np.max(np.abs(res)) <= tol:
            break
        if iters >= 100:
            raise RuntimeError("closure did not converge")
        J = -np.diag(par["k"] / hp + par["B"]) + Lmat @ np.diag(np.where(T0 < par["Tm"], phi, 0.0))
        T0 = T0 - np.linalg.solve(J, res)
        iters += 1
    Ti = jl_min(T0, par["Tm"])
    Ti = np.where(h == 0.0, 0.0, Ti)
    return T0, Ti, iters


def solveTi_trust_region(T0, x, t, h, Tw, phi, f, par, kind, cache, Lmat, abstol=1e-8, reltol=1e-6, norm="inf",
                         radius0=None, maxit=1000):
    """src/miz.jl:47-68 with a trust-region dogleg iteration (the algorithm family of NonlinearSolve.TrustRegion(),
    which is not vendored: Project.toml:33 compat "4.12.0", no Manifest).  Published algorithm restated: Nocedal &
    Wright, Numerical Optimization, Alg. 4.1 with the dogleg step (4.16) on m(p) = 1/2 |r + J p|^2, J the generalised
    Jacobian of the piecewise-linear residual; radius shrinks by 4 when rho < 1/4, doubles when rho > 3/4 on the
    boundary; a step is accepted when rho > 1e-4.  Stop as the reference asks (miz.jl:59): |r| <= abstol, or the
    step is below reltol * |T0| (NonlinearSolve's relative criterion), warm start = previous T0.
    Returns (T0, Ti, iterations, accepted_newton_steps)."""
    hp = np.where(h == 0.0, par["hmin"], h)
    nrm = (lambda v: float(np.max(np.abs(v)))) if norm == "inf" else (lambda v: float(np.sqrt(np.mean(v * v))))
    res = T0eq(T0, x, t, hp, Tw, phi, f, par, kind, cache)
    radius = radius0 if radius0 is not None else max(1.0, float(np.linalg.norm(T0)))
    it = newton_steps = 0
    while nrm(res) > abstol and it < maxit:
        J = -np.diag(par["k"] / hp + par["B"]) + Lmat @ np.diag(np.where(T0 < par["Tm"], phi, 0.0))
        g = J.T @ res
        pN = -np.linalg.solve(J, res)
        if np.linalg.norm(pN) <= radius:
            p, is_newton = pN, True
        else:
            Jg = J @ g
            pC = -(g @ g) / (Jg @ Jg) * g
            is_newton = False
            if np.linalg.norm(pC) >= radius:
                p = -radius / np.linalg.norm(g) * g
            else:                                   # dogleg: pC + tau (pN - pC) on the boundary
                d = pN - pC
                a_, b_, c_ = d @ d, 2 * (pC @ d), pC @ pC - radius**2
                tau = (-b_ + np.sqrt(b_ * b_ - 4 * a_ * c_)) / (2 * a_)
                p = pC + tau * d
        new = T0eq(T0 + p, x, t, hp, Tw, phi, f, par, kind, cache)
        pred = 0.5 * (res @ res) - 0.5 * np.sum((res + J @ p) ** 2)
        actual = 0.5 * (res @ res) - 0.5 * (new @ new)
        rho = actual / pred if pred > 0 else -1.0
        if rho < 0.25:
            radius *= 0.25
        elif rho > 0.75 and not is_newton:
            radius *= 2.0
        it += 1
        if rho > 1e-4:
            step_small = nrm(p) <= reltol * max(nrm(T0 + p), 1e-300) and nrm(new) <= abstol * 1e2
            T0, res = T0 + p, new
            newton_steps += is_newton
            if step_small:
                break
    Ti = jl_min(T0, par["Tm"])
    Ti = np.where(h == 0.0, 0.0, Ti)
    return T0, Ti, it, newton_steps


def miz_step(state, T0, x, t, f, dt, par, kind, cache, Lmat, tol=1e-8, closure="newton"):
    """src/miz.jl:150-196.  ``state`` has Ei, Ew, h, D, phi; returns (new_vars(10), T0, iters)."""
    Ei, Ew, h, D, phi = (state[k] for k in ("Ei", "Ew", "h", "D", "phi"))
    with np.errstate(all="ignore"):
        Tw = par["Tm"] + Ew / ((1 - phi) * par["cw"])
        Tw = np.where(np.isnan(Tw), 0.0, Tw)
        if closure == "newton":
            T0, Ti, iters = solveTi(T0, x, t, h, Tw, phi, f, par, kind, cache, Lmat, tol)
        else:
            T0, Ti, iters, _ = solveTi_trust_region(T0, x, t, h, Tw, phi, f, par, kind, cache, Lmat, abstol=tol,
                                                    norm="inf" if closure == "trust_region" else "rms")
        n = np.where(D == 0.0, 0.0, phi / (par["alpha"] * D**2))
        tb = Tbar(Ti, Tw, phi)
        L = par["A"] + par["B"] * (tb - par["Tm"])
        dif = diffusion(tb, x, par["D"], kind, cache)
        Fvi = solar(x, t, True, par) - L + dif + par["Fb"] + f
        Fvw = solar(x, t, False, par) - L + dif + par["Fb"] + f
        wl = par["m1"] * (Tw - par["Tm"] ** par["m2"])
        Flat = np.where(D == 0.0, 0.0, phi * h * par["Lf"] * wl * PI / (par["alpha"] * D))
        rEi = Ei + (phi * Fvi + Flat) * dt
        rEw = Ew + ((1 - phi) * Fvw - Flat) * dt
        cEi = np.where(rEi > 0.0, 0.0, rEi)
        cEw = np.where(rEw < 0.0, 0.0, rEw)
        psiEidt, psiEwdt = rEi - cEi, rEw - cEw
        Ei_n, Ew_n = cEi + psiEwdt, cEw + psiEidt
        Al = jl_min(par["alpha"] * n * ((D + 2.0 * par["rl"]) ** 2 - D**2), 1.0 - phi)
        psiEw = psiEwdt / dt
        Ql = np.where(phi == 1.0, 0.0, Al / (1 - phi) * psiEw)
        Qp = psiEw - Ql
        dn = dt * (-Qp / (par["Lf"] * par["alpha"] * par["Dmin"] ** 2 * par["hmin"]))
        lat_melt = -PI / 2.0 * par["alpha"] * wl
        lat_grow = np.where(h == 0.0, 0.0, -D / (2 * par["Lf"] * h * phi) * Ql)
        weld = par["kappa"] * par["alpha"] / 4 * phi * D**3
        rD = D + (lat_melt + lat_grow + weld) * dt
        total = n + dn
        Dn = np.where(total == 0.0, 0.0, (n * rD + dn * par["Dmin"]) / total)
        Dn = np.where(Dn > par["Dmax"], par["Dmax"], np.where(Dn < par["Dmin"], par["Dmin"], Dn))
        Dn = np.where(Ei_n == 0.0, 0.0, Dn)
        rh = h + (-1 / par["Lf"] * Fvi) * dt
        rh = np.where(rh < 0.0, 0.0, rh)
        hn = np.where(total == 0.0, 0.0, (n * rh + dn * par["hmin"]) / total)
        ph = np.where(hn == 0.0, 0.0, -Ei_n / (par["Lf"] * hn))
        ph = np.where(ph > 1.0, 1.0, ph)
        Ei_n = np.where(hn == 0.0, 0.0, Ei_n)
        E = ph * Ei_n + (1 - ph) * Ew_n
        T = Tbar(Ti, Tw, ph)
        Ti_out = np.where(Ei_n == 0.0, np.nan, Ti)
        Tw_out = np.where(ph > 0.99, np.nan, Tw)
    new = dict(T=T, Ei=Ei_n, Ti=Ti_out, D=Dn, n=n, h=hn, phi=ph, E=E, Ew=Ew_n, Tw=Tw_out)
    return new, T0, iters


def miz_integrate(st, forcing, par, init, nsteps=None, tol=1e-8, closure="newton", T0=None, start_step=0):
    """integrate(:MIZ, ...) storing every step; ``nsteps`` truncates the run (tests).  ``closure``: "newton"
    (semi-smooth Newton, what the oracle and the kernels use), "trust_region" / "trust_region_rms" (dogleg iteration
    with the reference's tolerances, max / rms residual norm).  ``T0`` / ``start_step``: continue from a given state."""
    kind = st.grid_kind
    cache = generic_stencil_cache(st.x) if kind == 1 else None
    Lmat = diffusion_matrix(st.x, par["D"], kind, cache)
    state = {k: np.array(init[k], dtype=float).copy() for k in ("Ei", "Ew", "h", "D", "phi")}
    T0 = np.zeros(st.nx) if T0 is None else np.array(T0, dtype=float).copy()
    n = st.nt * st.dur if nsteps is None else nsteps
    names = ("T", "Ei", "Ti", "D", "n", "h", "phi", "E", "Ew", "Tw")
    out = {k: np.empty((n, st.nx)) for k in names}
    total_iters = 0
    for q in range(1, n + 1):
        tinx = q + start_step
        ti = (tinx - 1) % st.nt + 1
        new, T0, iters = miz_step(state, T0, st.x, st.t[ti - 1], forcing(st.T(tinx)), st.dt, par, kind, cache, Lmat, tol, closure)
        total_iters += iters
        for k in names:
            out[k][q - 1] = new[k]
        state = {k: new[k] for k in ("Ei", "Ew", "h", "D", "phi")}
    out["_iters"] = total_iters
    out["_T0"] = T0
    return out
