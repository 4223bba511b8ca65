"""NumPy emulations of the parallel tridiagonal schemes the CUDA kernels use, checked against a dense solve.

These are the algorithms, not the kernels (the kernels are checked against the oracle on the GPU): partitioned
elimination with a left spike per band, the interface system solved by (a) parallel cyclic reduction over the lanes
of a warp (MIZ closure, csrc/miz_kernel.cu::tridiag) or (b) two sweeps from both ends meeting in the middle with
determinant-form pivots (classic, csrc/classic_uniform.cu), and the determinant form of the band pivots.
"""
import numpy as np
import pytest


def _system(rng, n, pad_to):
    """strictly diagonally dominant tridiagonal system of n rows, padded with decoupled identity-like rows"""
    jl = np.zeros(pad_to); jd = np.zeros(pad_to); ju = np.zeros(pad_to); rhs = np.zeros(pad_to)
    jl[1:n] = rng.uniform(0, 3000, n - 1); ju[:n - 1] = rng.uniform(0, 3000, n - 1)
    jd[:n] = -(22 + rng.uniform(0, 1, n)) - (jl[:n] + ju[:n])
    jd[n:] = -22.1
    rhs[:n] = rng.normal(size=n)
    A = np.diag(jd) + np.diag(jl[1:], -1) + np.diag(ju[:-1], 1)
    return jl, jd, ju, rhs, np.linalg.solve(A, rhs)


def _local_elimination(JL, JD, JU, R):
    """x_i + q_i x_{i+1} + s_i xL = y_i per band; returns q, s, y and (al, be, ga): x_0 = al - be*xL - ga*z"""
    nb, K = JD.shape
    q = np.zeros((nb, K)); s = np.zeros((nb, K)); y = np.zeros((nb, K))
    for i in range(K):
        w = JD[:, i] if i == 0 else JD[:, i] - JL[:, i] * q[:, i - 1]
        iw = 1.0 / w
        tq = JL[:, i] * iw
        q[:, i] = JU[:, i] * iw
        y[:, i] = R[:, i] * iw if i == 0 else R[:, i] * iw - tq * y[:, i - 1]
        s[:, i] = tq if i == 0 else -tq * s[:, i - 1]
    al, be, ga = y[:, K - 2].copy(), s[:, K - 2].copy(), q[:, K - 2].copy()
    for i in range(K - 3, -1, -1):
        al = y[:, i] - q[:, i] * al; be = s[:, i] - q[:, i] * be; ga = -q[:, i] * ga
    return q, s, y, al, be, ga


def _interface_rows(q, s, y, al, be, ga):
    nb, K = q.shape
    nxt = lambda v: np.r_[v[1:], 0.0]
    ql = q[:, K - 1].copy(); ql[-1] = 0.0
    A = s[:, K - 1].copy(); A[0] = 0.0
    B = 1.0 - ql * nxt(be); C = -ql * nxt(ga); R = y[:, K - 1] - ql * nxt(al)
    return A / B, C / B, R / B


def _back_substitute(q, s, y, z):
    nb, K = q.shape
    xL = np.r_[0.0, z[:-1]]
    X = np.zeros((nb, K)); X[:, K - 1] = z
    xn = z.copy()
    for i in range(K - 2, -1, -1):
        xn = y[:, i] - s[:, i] * xL - q[:, i] * xn
        X[:, i] = xn
    return X


@pytest.mark.parametrize("n,K", [(180, 6), (100, 4), (250, 8), (50, 2)])
def test_partitioned_elimination_with_pcr_interface(n, K):
    """MIZ: 32 lanes x K rows, PCR (5 steps, rows normalised to unit diagonal, out-of-range neighbours = identity)."""
    rng = np.random.default_rng(n)
    jl, jd, ju, rhs, xref = _system(rng, n, 32 * K)
    q, s, y, al, be, ga = _local_elimination(jl.reshape(32, K), jd.reshape(32, K), ju.reshape(32, K), rhs.reshape(32, K))
    A, C, R = _interface_rows(q, s, y, al, be, ga)
    lane = np.arange(32)
    st = 1
    while st < 32:
        up = lambda v: np.r_[v[:st], v[:-st]]
        dn = lambda v: np.r_[v[st:], v[-st:]]
        a_ = np.where(lane >= st, A, 0.0); c_ = np.where(lane + st < 32, C, 0.0)
        ib = 1.0 / (1.0 - a_ * up(C) - c_ * dn(A))
        Rn = (R - a_ * up(R) - c_ * dn(R)) * ib
        A, C, R = -(a_ * up(A)) * ib, -(c_ * dn(C)) * ib, Rn
        st *= 2
    X = _back_substitute(q, s, y, R)
    assert np.abs(X.reshape(-1) - xref).max() < 1e-13 * max(1.0, np.abs(xref).max())
    assert np.abs(A).max() < 1e-19 and np.abs(C).max() < 1e-19     # all couplings are gone after log2(32) steps


def test_partitioned_elimination_with_two_ended_interface():
    """Classic: 8 bands x 13 rows; the 8 interface unknowns are eliminated from both ends with determinant-form
    pivots D_k = d_k D_{k-1} - a_k c_{k-1} D_{k-2} and the two sweeps meet between rows 3 and 4."""
    rng = np.random.default_rng(7)
    K, WB = 13, 8
    jl, jd, ju, rhs, xref = _system(rng, 100, WB * K)
    jd = -jd; jl = -jl; ju = -ju; rhs = -rhs                      # the classic matrix has a positive diagonal
    q, s, y, al, be, ga = _local_elimination(jl.reshape(WB, K), jd.reshape(WB, K), ju.reshape(WB, K), rhs.reshape(WB, K))
    nxt = lambda v: np.r_[v[1:], 0.0]
    ql = q[:, K - 1]
    dg = 1.0 - ql * nxt(be); sup = -ql * nxt(ga); r = y[:, K - 1] - ql * nxt(al); sl = s[:, K - 1].copy(); sl[0] = 0.0
    H = WB // 2
    z = np.zeros(WB)
    sweeps = []
    for half in (0, 1):
        rows = [WB - 1 - k if half else k for k in range(H)]
        a_ = [sup[b] if half else sl[b] for b in rows]           # coupling to the previously eliminated row
        c_ = [sl[b] if half else sup[b] for b in rows]           # coupling to the next row in sweep order
        Dm = [1.0, dg[rows[0]]]
        for k in range(1, H):
            Dm.append(dg[rows[k]] * Dm[k] - (a_[k] * c_[k - 1]) * Dm[k - 1])
        cq, cy = [], []
        for k in range(H):
            iw = Dm[k] / Dm[k + 1]
            cq.append(c_[k] * iw)
            cy.append(r[rows[k]] * iw - (a_[k] * iw) * (cy[k - 1] if k else 0.0))
        sweeps.append((rows, cq, cy))
    (r0, cq0, cy0), (r1, cq1, cy1) = sweeps
    x0 = (cy0[-1] - cq0[-1] * cy1[-1]) / (1.0 - cq0[-1] * cq1[-1])
    x1 = (cy1[-1] - cq1[-1] * cy0[-1]) / (1.0 - cq1[-1] * cq0[-1])
    for (rows, cq, cy), x in ((sweeps[0], x0), (sweeps[1], x1)):
        z[rows[-1]] = x
        for k in range(H - 2, -1, -1):
            x = cy[k] - cq[k] * x
            z[rows[k]] = x
    X = _back_substitute(q, s, y, z)
    assert np.abs(X.reshape(-1)[:100] - xref[:100]).max() < 1e-13 * max(1.0, np.abs(xref).max())


def test_determinant_form_pivots_equal_thomas_pivots():
    """1/w_i = P_{i-1}/P_i with P_i = d_i P_{i-1} - (a_i c_{i-1}) P_{i-2}: one dependent FMA per row, reciprocals
    independent of each other (classic masked bands); no overflow for 13 rows of |w| <= 250."""
    rng = np.random.default_rng(11)
    K = 13
    d = rng.uniform(51, 250, K); off = -rng.uniform(0, 30, K)
    a = off.copy(); a[0] = 0.0; c = np.r_[off[1:], 0.0]
    w = np.zeros(K); w[0] = d[0]
    for i in range(1, K):
        w[i] = d[i] - a[i] * c[i - 1] / w[i - 1]
    P = [1.0, d[0]]
    for i in range(1, K):
        P.append(d[i] * P[i] - (a[i] * c[i - 1]) * P[i - 1])
    iw = np.array([P[i] / P[i + 1] for i in range(K)])
    assert np.abs(iw * w - 1.0).max() < 1e-14 and np.isfinite(P[-1]) and abs(P[-1]) < 1e40


def _local_elimination_determinant(JL, JD, JU, R):
    """csrc/miz_kernel.cu::tridiag, round-2 form: pivots through P_i = d_i P_{i-1} - (l_i u_{i-1}) P_{i-2} (one dependent FMA
    per row; the K reciprocals independent), rows scaled first, then the y / spike chains."""
    nb, K = JD.shape
    q = np.zeros((nb, K)); s = np.zeros((nb, K)); y = np.zeros((nb, K))
    Pm2 = np.ones(nb); Pm1 = np.ones(nb); jup = np.zeros(nb); Pmax = 0.0
    for i in range(K):
        P = JD[:, i] if i == 0 else JD[:, i] * Pm1 - (JL[:, i] * jup) * Pm2
        iw = Pm1 / P
        s[:, i] = JL[:, i] * iw; q[:, i] = JU[:, i] * iw; y[:, i] = R[:, i] * iw
        Pm2, Pm1, jup = Pm1, P, JU[:, i]
        Pmax = max(Pmax, np.abs(P).max())
    for i in range(1, K):
        y[:, i] = y[:, i] - s[:, i] * y[:, i - 1]
        s[:, i] = -s[:, i] * s[:, i - 1]
    al, be, ga = y[:, K - 2].copy(), s[:, K - 2].copy(), q[:, K - 2].copy()
    for i in range(K - 3, -1, -1):
        al = y[:, i] - q[:, i] * al; be = s[:, i] - q[:, i] * be; ga = -q[:, i] * ga
    return q, s, y, al, be, ga, Pmax


@pytest.mark.parametrize("n,K", [(180, 6), (250, 8), (100, 4)])
def test_miz_band_elimination_in_determinant_form(n, K):
    """The determinant-form band elimination of the MIZ closure equals the textbook one (same q, s, y to rounding) and
    its determinants stay far inside the double range at the magnitudes of the MIZ Jacobian (|w| up to ~4e4, K <= 8);
    the PCR interface may stop as soon as the couplings are below 1e-19, tested from stride 4 on as in the kernel."""
    rng = np.random.default_rng(7 * n + K)
    jl, jd, ju, rhs, xref = _system(rng, n, 32 * K)
    jl *= 6.0; ju *= 6.0                                   # couplings up to 1.8e4 (D = 0.6, nx = 180: 0.6 / dx^2)
    jd[:n] = -(22 + rng.uniform(0, 1, n)) - (jl[:n] + ju[:n])
    # open-water stretch: decoupled rows (g = 0), as for a member with little ice
    jl[40:n] = 0.0; ju[40:n] = 0.0; jd[40:n] = -22.1
    A = np.diag(jd) + np.diag(jl[1:], -1) + np.diag(ju[:-1], 1)
    xref = np.linalg.solve(A, rhs)
    shp = (32, K)
    q0, s0, y0, *_ = _local_elimination(jl.reshape(shp), jd.reshape(shp), ju.reshape(shp), rhs.reshape(shp))
    q, s, y, al, be, ga, Pmax = _local_elimination_determinant(jl.reshape(shp), jd.reshape(shp), ju.reshape(shp), rhs.reshape(shp))
    assert Pmax < 1e60
    for a_, b_ in ((q, q0), (s, s0), (y, y0)):
        assert np.abs(a_ - b_).max() <= 1e-12 * max(1.0, np.abs(b_).max())
    Ai, Ci, Ri = _interface_rows(q, s, y, al, be, ga)
    lane = np.arange(32)
    st, steps = 1, 0
    while st < 32:
        if st >= 4 and not (np.abs(Ai) + np.abs(Ci) > 1e-19).any():
            break
        up = lambda v: np.r_[v[:st], v[:-st]]
        dn = lambda v: np.r_[v[st:], v[-st:]]
        a_ = np.where(lane >= st, Ai, 0.0); c_ = np.where(lane + st < 32, Ci, 0.0)
        ib = 1.0 / (1.0 - a_ * up(Ci) - c_ * dn(Ai))
        Rn = (Ri - a_ * up(Ri) - c_ * dn(Ri)) * ib
        Ai, Ci, Ri = -(a_ * up(Ai)) * ib, -(c_ * dn(Ci)) * ib, Rn
        st *= 2; steps += 1
    X = _back_substitute(q, s, y, Ri)
    assert np.abs(X.reshape(-1) - xref).max() < 1e-12 * max(1.0, np.abs(xref).max())
    assert steps <= 5


def test_pad_rows_do_not_reach_the_real_cells():
    """Classic, 100 cells on 8 bands of 13: the 4 pad rows of the last band are decoupled (zero off-diagonals), so
    whatever their diagonal and right-hand side -- the state the pad cells happen to be in -- the solution of the 100
    real rows is the same to the last bit of the elimination's arithmetic."""
    rng = np.random.default_rng(5)
    n, K, nb = 100, 13, 8
    sols = []
    for trial in range(3):
        r = np.random.default_rng(99)                       # the same real system every trial
        jl = np.zeros(nb * K); jd = np.zeros(nb * K); ju = np.zeros(nb * K); rhs = np.zeros(nb * K)
        jl[1:n] = -r.uniform(0, 30, n - 1); ju[:n - 1] = jl[1:n]
        jd[:n] = 51.0 + r.uniform(0, 60, n) - r.uniform(0, 49, n) * (r.uniform(size=n) < 0.5)   # masked rows lose up to dc/g
        rhs[:n] = r.normal(size=n) * 10
        jd[n:] = rng.uniform(2, 60, nb * K - n)             # pad rows: different every trial
        rhs[n:] = rng.normal(size=nb * K - n) * 1e3
        shp = (nb, K)
        q, s, y, al, be, ga = _local_elimination(jl.reshape(shp), jd.reshape(shp), ju.reshape(shp), rhs.reshape(shp))
        Ai, Ci, Ri = _interface_rows(q, s, y, al, be, ga)
        T = np.diag(np.ones(nb)) + np.diag(Ai[1:], -1) + np.diag(Ci[:-1], 1)
        z = np.linalg.solve(T, Ri)
        sols.append(_back_substitute(q, s, y, z).reshape(-1)[:n])
        A = np.diag(jd) + np.diag(jl[1:], -1) + np.diag(ju[:-1], 1)
        assert np.abs(sols[-1] - np.linalg.solve(A, rhs)[:n]).max() < 1e-11
    # bands 0..6 never see the pad rows; in band 7 the pad rows sit after the real ones, so its forward pass is untouched
    assert np.array_equal(sols[0][:91], sols[1][:91]) and np.array_equal(sols[0][:91], sols[2][:91])
    assert np.abs(sols[0] - sols[1]).max() < 1e-12 and np.abs(sols[0] - sols[2]).max() < 1e-12
