"""CPU tests of the oracle (the checker): against the reference's docstring known-answer values, against an
independent NumPy restatement, against the committed golden vectors, and through invariants of the model.

PARITY UNPINNED: the reference's only fixture (test/solution_1year.jld2) is absent and Julia is not installed; the
golden vectors under tests/golden/ were produced by this same oracle (scripts/make_golden.py).
"""
import os

import numpy as np
import pytest

import ebm_b200 as ebm
import np_restatement as npr
import oracle
from helpers import cold_init, oracle_classic, oracle_diag_classic, oracle_miz, rel_err, warm_init

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _zero(nx):
    z = np.zeros(nx)
    return ebm.Collection(Ei=z.copy(), Ew=z.copy(), h=z.copy(), D=z.copy(), phi=z.copy())


# ---------------------------------------------------------------- docstring known answers
def test_spacetime_known_answers():
    st = ebm.SpaceTime(180, 2000, 30, "sin")                     # src/infrastructure.jl:101-106
    assert abs(st.x[0] - 0.00436331) < 5e-9 and abs(st.x[1] - 0.0130896) < 5e-8
    assert abs(st.x[-2] - 0.999914) < 5e-7 and abs(st.x[-1] - 0.99999) < 5e-6
    assert st.t[0] == 0.00025 and st.t[1] == 0.00075 and st.t[-1] == 0.99975   # :97
    assert st.winter.inx == 522 and st.summer.inx == 1548        # round-half-even of 522.5 / 1547.5
    assert st.grid_kind == 1 and ebm.SpaceTime(100, 2000, 1).grid_kind == 0
    sols = ebm.Solutions(st, ebm.Forcing(0.0), {}, {}, ebm.MIZ_VARS, True)     # src/EnergyBalanceModel.jl:65
    assert abs(sols.ts[0] - 29.00025) < 1e-12 and abs(sols.ts[-1] - 29.99975) < 1e-12 and len(sols.ts) == 2000


def test_forcing_known_answers_and_errors():
    f = ebm.Forcing(0.0, 5.0, -5.0, (10, 10), (0.5, -0.5))       # src/infrastructure.jl:193-205
    assert f.domain == (0, 10, 20, 30, 50)
    assert abs(f(17.57) - 3.785) < 1e-12
    assert f(5.0) == 0.0 and f(25.0) == 5.0 and f(60.0) == -5.0
    assert abs(oracle.forcing(f.row(), 17.57) - f(17.57)) == 0.0
    assert oracle.forcing(ebm.Forcing(2.5).row(), 123.4) == 2.5
    with pytest.raises(ValueError):
        ebm.Forcing(0.0, 5.0, -5.0, (10, 10), (0.3, -0.5))       # warming time not an integer (:231)
    with pytest.raises(ValueError):
        ebm.Forcing(0.0, 5.0, -5.0, (10, 10), (0.5, 0.5))        # cooling rate must be negative (:238)


def test_forcing_annual_mean_and_hysteresis_points():
    f = ebm.Forcing(0.0, 5.0, -5.0, (10, 10), (0.5, -0.5))
    st = ebm.SpaceTime(100, 2000, 60)
    assert ebm.annual_mean_forcing(f, st, 5) == 0.0                          # hold at base
    assert abs(ebm.annual_mean_forcing(f, st, 11) - 0.25) < 1e-12           # first warming year: mean of 0..0.5
    assert abs(ebm.annual_mean_forcing(f, st, 25) - 5.0) < 1e-12            # hold at peak
    assert ebm.annual_mean_forcing(ebm.Forcing(1.5), st, 3) == 1.5
    # plot_seasonal's x / y functions (src/plot.jl:173-190) on oracle fields == the L0 diagnostics layout
    st1 = ebm.SpaceTime(100, 2000, 1)
    o = oracle_classic(st1, [ebm.Forcing(0.0)], [ebm.default_parameters("Classic")], [cold_init(100)], seasonal=True)
    d = oracle_diag_classic(o["seasonal"], st1.x)
    x, y = ebm.hysteresis_points(d, season=0)
    assert abs(x[0, 0] - ebm.hemispheric_mean(o["seasonal"][0, 0, 2, 1], st1.x)) < 1e-12
    ice = (o["seasonal"][0, 0, 0, 0] < 0).astype(float)
    assert abs(y[0, 0] - 2 * np.pi * ebm.hemispheric_mean(ice, st1.x)) < 1e-12


def test_default_parameters():
    pm, pc = ebm.default_parameters("MIZ"), ebm.default_parameters("Classic")
    assert len(pm) == 22 and len(pc) == 16                       # src/EnergyBalanceModel.jl:29-41, infrastructure.jl:458
    assert abs(pm.m1 - 50.4576) < 1e-12 and pm.kappa == 315360.0 and abs(pc.cg - 0.098) < 1e-15
    assert tuple(ebm.MIZ_PAR_ORDER) == ebm.miz_paramset and "F" not in ebm.CLASSIC_PAR_ORDER
    with pytest.raises(ValueError):
        ebm.model_name(":classic")                               # the reference dispatches on Val{:Classic} only


# ---------------------------------------------------------------- oracle vs the independent NumPy restatement
def test_classic_oracle_matches_numpy_restatement():
    st = ebm.SpaceTime(100, 2000, 1)
    par = ebm.default_parameters("Classic")
    for init, F in ((warm_init(100), 0.0), (cold_init(100), 2.0)):
        o = oracle_classic(st, [ebm.Forcing(F)], [par], [init], lastonly=False, raw=True)
        r = npr.classic_integrate(ebm.SpaceTime(100, 2000, 1), ebm.Forcing(F), par, init.E, init.Tg)
        n = 300   # the restatement uses a dense solve: compare the first 300 steps
        for vi, v in enumerate(("E", "T", "h")):
            assert rel_err(o["raw"][0, :n, vi], r[v][:n]).max() < 1e-10, v


def test_classic_dense_lu_variant_bounds_solver_divergence():
    """The reference factorises a dense matrix (classic.jl:55-63); the oracle's tridiagonal solve must agree."""
    st = ebm.SpaceTime(100, 2000, 2)
    par = ebm.default_parameters("Classic")
    a = oracle_classic(st, [ebm.Forcing(0.0)], [par], [warm_init(100)], solver=oracle.SOLVE_TRIDIAG)
    b = oracle_classic(st, [ebm.Forcing(0.0)], [par], [warm_init(100)], solver=oracle.SOLVE_DENSE_LU)
    assert rel_err(a["E"], b["E"]).max() < 1e-10 and rel_err(a["Tg"], b["Tg"]).max() < 1e-10


@pytest.mark.parametrize("xfunc,nx", [("sin", 180), ("identity", 60)])
def test_miz_oracle_matches_numpy_restatement(xfunc, nx):
    st = ebm.SpaceTime(nx, 2000, 1, xfunc)
    par = ebm.default_parameters("MIZ")
    o = oracle_miz(st, [ebm.Forcing(0.0)], [par], [_zero(nx)], lastonly=False, raw=True)
    n = 15
    r = npr.miz_integrate(st, ebm.Forcing(0.0), par, _zero(nx), nsteps=n)
    for vi, v in enumerate(ebm.MIZ_VARS):
        a, b = o["raw"][0, :n, vi], r[v]
        assert np.array_equal(np.isnan(a), np.isnan(b)), v
        assert rel_err(a, b).max() < 1.5e-8, (v, rel_err(a, b).max())


# ---------------------------------------------------------------- golden vectors
def test_miz_golden_fixture_setup():
    g = np.load(os.path.join(GOLD, "miz_fixture_setup.npz"))
    st = ebm.SpaceTime(180, 2000, 1, "sin")
    o = oracle_miz(st, [ebm.Forcing(0.0)], [ebm.default_parameters("MIZ")], [_zero(180)], lastonly=False, raw=True)
    assert tuple(g["variables"]) == ebm.MIZ_VARS
    # the reference's criterion (test/runtests.jl:40-46): NaN -> 0, isapprox(rtol = sqrt(eps))
    a, b = np.nan_to_num(o["raw"][0, 9]), np.nan_to_num(g["step10"])
    assert np.all(np.abs(a - b) <= 1.4901161193847656e-08 * np.maximum(np.abs(a), np.abs(b)))
    assert np.array_equal(o["raw"][0, :20], g["first20"], equal_nan=True)   # same build of the oracle: bitwise
    # SURVEY Appendix D probe values at step 10
    E, Ti = g["step10"][ebm.MIZ_VARS.index("E")], np.nan_to_num(g["step10"][ebm.MIZ_VARS.index("Ti")])
    assert abs(E[0] - 0.509) < 1e-3 and abs(E[-1] + 1.174) < 1e-3 and abs(Ti[-1] + 11.24) < 1e-2
    assert int((g["step10"][ebm.MIZ_VARS.index("phi")] > 0).sum()) == 149


def test_classic_golden_and_climate():
    g = np.load(os.path.join(GOLD, "classic_default.npz"))
    st = ebm.SpaceTime(100, 2000, 1)
    par = ebm.default_parameters("Classic")
    o = oracle_classic(st, [ebm.Forcing(0.0)], [par], [warm_init(100)], lastonly=False, raw=True)
    assert rel_err(o["raw"][0, 99::100], g["every100"]).max() < 1e-12
    assert rel_err(o["E"][0], g["final_E"]).max() < 1e-12 and rel_err(o["Tg"][0], g["final_Tg"]).max() < 1e-12
    # year-30 climate of the warm branch: hemispheric-mean annual T ~ 16.98, seasonal ice edge
    seas = g["y30_seasonal"]                                       # [2 members][3 seasons][3 vars][nx]
    x = ebm.SpaceTime(100, 2000, 30).x
    hm = ebm.hemispheric_mean(seas[0, 2, 1], x)
    assert 16.9 < hm < 17.1
    d = oracle_diag_classic(seas[:, None], x)                      # [2, 1, 3, 4]
    assert 0.80 < d[0, 0, 0, 3] < 0.87 and 0.95 < d[0, 0, 1, 3] <= 1.0     # winter / summer ice edge
    assert d[1, 0, 2, 2] > 6.2                                     # cold start: snowball (ice area ~ 2*pi)


# ---------------------------------------------------------------- invariants
def test_diffusion_operators_conserve_and_agree():
    """sum_j diffusion_j * cell width_j = 0 (no-flux ends) for both operators; on the identity grid the generic
    stencil equals get_diffop."""
    rng = np.random.default_rng(1)
    for xfunc, nx in (("identity", 100), ("sin", 180)):
        st = ebm.SpaceTime(nx, 10, 1, xfunc)
        T = rng.normal(size=nx)
        cache = npr.generic_stencil_cache(st.x)
        d = npr.diffusion(T, st.x, 0.6, 1, cache)
        width = cache[3] if isinstance(cache, (tuple, list)) else None
        xe = np.concatenate([[-st.x[0]], st.x, [2 - st.x[-1]]])
        w = (xe[2:] + xe[1:-1]) / 2 - (xe[1:-1] + xe[:-2]) / 2
        assert abs(np.sum(d * w)) < 1e-9 * np.abs(d * w).sum()
        if xfunc == "identity":
            d0 = npr.diffusion(T, st.x, 0.6, 0)
            assert np.abs(d - d0).max() < 1e-10 * np.abs(d0).max()


def test_classic_energy_budget_and_sampling_semantics():
    """E_{n+1} - E_n = dt*(C - M*T + Fb) is what the step does; here: h = -E/Lf*(E<0), T = E_prev/cw where E_prev >= 0,
    winter/summer snapshots are raw steps 522/1548, the annual mean is the mean over the year (savesol!)."""
    st = ebm.SpaceTime(100, 2000, 2)
    par = ebm.default_parameters("Classic")
    o = oracle_classic(st, [ebm.Forcing(1.0)], [par], [warm_init(100)], lastonly=False, raw=True, seasonal=True)
    E, T, h = o["raw"][0, :, 0], o["raw"][0, :, 1], o["raw"][0, :, 2]
    assert np.array_equal(h, np.where(E < 0, -E / par.Lf, 0.0))
    Ep, Tn = E[:-1], T[1:]
    assert np.array_equal(Tn[Ep >= 0], Ep[Ep >= 0] / par.cw)
    for y in range(2):
        assert np.array_equal(o["seasonal"][0, y, 0], o["raw"][0, y * 2000 + 521])
        assert np.array_equal(o["seasonal"][0, y, 1], o["raw"][0, y * 2000 + 1547])
        assert rel_err(o["seasonal"][0, y, 2], o["raw"][0, y * 2000:(y + 1) * 2000].mean(axis=0)).max() < 1e-12


def test_miz_state_invariants_and_closure_statistics():
    st = ebm.SpaceTime(180, 2000, 2, "sin")
    par = ebm.default_parameters("MIZ")
    o = oracle_miz(st, [ebm.Forcing(0.0)], [par], [_zero(180)], raw=True)
    assert (o["phi"] >= 0).all() and (o["phi"] <= 1).all() and (o["h"] >= 0).all()
    assert (o["Ei"] <= 0).all() and (o["Ew"] >= 0).all()
    D = o["D"]
    assert ((D == 0) | ((D >= par.Dmin) & (D <= par.Dmax))).all()
    assert o["nonconv"][0] == 0 and 1.0 <= o["newton_iters"][0] / 4000 <= 1.3    # SURVEY: 1.05-1.14 iterations/step
    Ti, Ei = o["raw"][0, :, ebm.MIZ_VARS.index("Ti")], o["raw"][0, :, ebm.MIZ_VARS.index("Ei")]
    assert np.array_equal(np.isnan(Ti), Ei == 0.0)                               # miz.jl:193


def test_miz_is_sensitive_to_rounding_level_perturbations():
    """Documents why long-run pointwise MIZ parity is undefined: 1e-13 on the initial state becomes
    O(0.1) within a few hundred steps of the spin-up, for the oracle itself."""
    st = ebm.SpaceTime(180, 2000, 1, "sin")
    par = ebm.default_parameters("MIZ")
    a = oracle_miz(st, [ebm.Forcing(0.0)], [par], [_zero(180)], lastonly=False, raw=True)
    p = _zero(180)
    p.Ew = p.Ew + 1e-13
    b = oracle_miz(st, [ebm.Forcing(0.0)], [par], [p], lastonly=False, raw=True)
    early = max(rel_err(b["raw"][0, 9, vi], a["raw"][0, 9, vi]).max() for vi in range(10))
    late = max(rel_err(b["raw"][0, 400:, vi], a["raw"][0, 400:, vi]).max() for vi in range(10))
    assert early < 1e-6 and late > 1e-3


# ---------------------------------------------------------------- the NonlinearSolve boundary (miz.jl:55-60)
@pytest.mark.parametrize("closure", ["trust_region", "trust_region_rms"])
def test_trust_region_closure_agrees_with_semismooth_newton(closure):
    """The reference solves the surface-temperature closure with NonlinearSolve.TrustRegion() (abstol 1e-8, reltol
    1e-6, warm start) -- a third-party dependency that is not vendored and whose version is unpinned.  The oracle and
    the kernels use a semi-smooth Newton iteration on the same residual.  This bounds the substitution: a trust-region
    dogleg iteration (published algorithm, Nocedal & Wright Alg. 4.1) with the reference's stop, run on the fixture
    setup (test/runtests.jl:22-32, steps 1-20) and from year-1 / year-10 / year-30 states of the docstring run (C3),
    gives the ten stored variables within the reference's own rtol = sqrt(eps) of the Newton run at step 10 and at
    step 20.  (The residual is piecewise linear with a tridiagonal generalised Jacobian: whenever the Newton step
    fits the trust radius both methods take the same step.)"""
    rtol = 1.4901161193847656e-08
    par = ebm.default_parameters("MIZ")
    f = ebm.Forcing(0.0)

    def compare(st, init, T0, start_step, what):
        a = npr.miz_integrate(st, f, par, init, nsteps=20, T0=T0, start_step=start_step)
        b = npr.miz_integrate(st, f, par, init, nsteps=20, T0=T0, start_step=start_step, closure=closure)
        for step in (9, 19):
            for v in ebm.MIZ_VARS:
                x, y = np.nan_to_num(a[v][step]), np.nan_to_num(b[v][step])
                assert np.array_equal(np.isnan(a[v][step]), np.isnan(b[v][step])), (what, v)
                assert np.all(np.abs(x - y) <= rtol * np.maximum(np.abs(x), np.abs(y))), (what, v, step, np.abs(x - y).max())
        return a["_iters"], b["_iters"]

    st1 = ebm.SpaceTime(180, 2000, 1, "sin")
    it_n, it_t = compare(st1, _zero(180), None, 0, "fixture setup")
    assert it_n >= 20 and it_t >= 20
    for years in (1, 10, 30):
        st = ebm.SpaceTime(180, 2000, years, "sin")
        o = oracle_miz(st, [f], [par], [_zero(180)])
        init = ebm.Collection(**{k: o[k][0] for k in ("Ei", "Ew", "h", "D", "phi")})
        compare(st, init, o["T0"][0], years * 2000, f"C3 after {years} years")


def test_c5_blow_up_members_blow_up_in_the_independent_restatement_too():
    """Part of the C5 sweep (D, B, ai, k, m1 over 16^5 values) has no finite solution in the reference
    ALGORITHM: lateral melt too weak (m1 below its default) lets the ice concentration reach phi = 1 next to warm water
    and water_temp (miz.jl:30) divides by 1 - phi = 0.  The C oracle and the independent NumPy restatement (different
    code, dense solves) agree on which members blow up and on the year; a neighbouring member with default m1 stays
    finite in both.  bench.py reports such members (`nan_members`) and the throughput over finite members."""
    import bench
    N = 16 ** 5
    # (member index in the 16^5 grid, blows up within 5 years?) -- found with the oracle on a 512-member sample
    cases = [(20776, True), (196911, True)]
    st5 = ebm.SpaceTime(180, 2000, 5, "sin")
    idx = np.array([c[0] for c in cases] + [0])
    _, rows, forc, init = bench.miz_workload(ebm, N, idx, 5)
    rows[-1] = [ebm.default_parameters("MIZ")[k] for k in ebm.MIZ_PAR_ORDER]        # default parameters: stable
    o = oracle.miz_run(st5.x, st5.t, 5, st5.winter.inx, st5.summer.inx, st5.grid_kind, rows, forc, *init)
    finite = np.isfinite(o["Ei"]).all(axis=1) & np.isfinite(o["Ew"]).all(axis=1)
    assert list(finite) == [False, False, True]
    order = list(ebm.MIZ_PAR_ORDER)
    for k, (m, blows) in enumerate(cases[:1] + [(None, False)]):
        par = ebm.Collection(dict(zip(order, rows[k if m is not None else -1])))
        blew = False
        try:
            with np.errstate(all="ignore"):
                r = npr.miz_integrate(st5, ebm.Forcing(0.0), par, _zero(180), nsteps=(5 if blows else 1) * 2000)
            blew = not (np.isfinite(r["Ei"][-1]).all() and np.isfinite(r["Ew"][-1]).all())
        except (RuntimeError, np.linalg.LinAlgError):
            blew = True                                   # closure residual is NaN: no convergence, or a singular Jacobian
        assert blew == blows, (m, blows)
    assert rows[0][order.index("m1")] > 0 and rows[1][order.index("ai")] < 0.4


def test_classic_generic_stencil_extension():
    """SURVEY 8f-4 (an extension: the reference's classic model uses get_diffop(nx) whatever the grid, classic.jl:21):
    kappa from the generic flux-form stencil (infrastructure.jl:500-527).  On the identity grid the two operators agree
    to rounding, so the runs agree far inside the tolerance; on a sin grid the oracle (tridiagonal solve with distinct
    sub/super diagonals) agrees with its dense-LU variant and with the independent NumPy restatement."""
    par = ebm.default_parameters("Classic")
    f = ebm.Forcing(0.0)
    st = ebm.SpaceTime(100, 2000, 1)
    a = oracle_classic(st, [f], [par], [warm_init(100)])
    b = oracle_classic(st, [f], [par], [warm_init(100)], stencil=1)
    assert rel_err(b["E"], a["E"]).max() < 1e-10 and rel_err(b["Tg"], a["Tg"]).max() < 1e-10
    sts = ebm.SpaceTime(60, 1000, 1, "sin")
    c = oracle_classic(sts, [f], [par], [warm_init(60)], stencil=1, lastonly=False, raw=True)
    d = oracle_classic(sts, [f], [par], [warm_init(60)], stencil=1, solver=oracle.SOLVE_DENSE_LU)
    assert rel_err(d["E"], c["E"]).max() < 1e-10
    r = npr.classic_integrate(sts, f, par, np.full(60, 98.0), np.full(60, 10.0), stencil=1)
    for vi, v in enumerate(("E", "T", "h")):
        assert rel_err(r[v], c["raw"][0, :, vi]).max() < 1e-9, v
    ref_behaviour = oracle_classic(sts, [f], [par], [warm_init(60)])          # get_diffop on the sin grid: what the reference does
    assert np.abs(ref_behaviour["E"] - c["E"]).max() > 1.0                    # ... a different model


def test_classic_step_debug_menu_matches_the_numpy_restatement():
    """The oracle's debug menu (locals of step!, src/classic.jl:47-56, that a `debug::Expr` would name) against the
    independent NumPy restatement, warm / cold / mixed states incl. E == 0."""
    import np_restatement as npr
    st = ebm.SpaceTime(100, 2000, 1)
    par = ebm.default_parameters("Classic")
    row = [par[k] for k in ebm.CLASSIC_PAR_ORDER]
    stat = npr.classic_statics(st.x, st.t, st.nt, par)
    E0 = np.linspace(60.0, -25.0, 100)
    E0[40] = 0.0
    Tg0 = np.linspace(12.0, -14.0, 100)
    for i1, f in ((1, 0.0), (7, 1.25), (2000, -3.0)):
        for name in oracle.DEBUG_MENU:
            o = oracle.classic_step(st.x, st.t, row, i1, f, E0, Tg0, debug=name)
            E, Tg, T, h, dbg = npr.classic_step(stat, par, i1, f, E0.copy(), Tg0.copy(), debug=name)
            np.testing.assert_allclose(o["debug"], dbg, rtol=1e-13, atol=1e-13, err_msg=name)
            np.testing.assert_allclose(o["E"], E, rtol=1e-13, atol=1e-12)
            np.testing.assert_allclose(o["Tg"], Tg, rtol=1e-11, atol=1e-11)
    o = oracle.classic_step(st.x, st.t, row, 7, 1.25, E0, Tg0, debug="alpha")
    assert o["debug"][40] == 0.0 and set(np.unique(o["debug"][E0 < 0.0])) == {par["ai"]}
