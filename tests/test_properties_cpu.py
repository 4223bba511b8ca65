"""Property tests (hypothesis) of the host mirror against the C oracle's helpers: Forcing, grids, means, sharding."""
import math

import numpy as np
from hypothesis import given, settings, strategies as st

import ebm_b200 as ebm
import oracle


@settings(max_examples=60, deadline=None)
@given(base=st.integers(-80, 80).map(lambda k: k / 4.0), up=st.integers(1, 20), down=st.integers(1, 20), h0=st.integers(0, 15), h1=st.integers(0, 15),
       ru=st.sampled_from([0.25, 0.5, 1.0, 2.0]), rd=st.sampled_from([-0.25, -0.5, -1.0, -2.0]), T=st.floats(0, 120))
def test_forcing_call_matches_oracle_and_is_continuous(base, up, down, h0, h1, ru, rd, T):
    peak, cool = base + ru * up, base + ru * up + rd * down
    f = ebm.Forcing(base, peak, cool, (h0, h1), (ru, rd))
    assert f.domain == (0, h0, h0 + up, h0 + up + h1, h0 + up + h1 + down)          # infrastructure.jl:221-240
    assert f(T) == oracle.forcing(f.row(), T)                                        # same branch chain, same arithmetic
    for d in f.domain[1:]:                                                           # piecewise linear and continuous
        assert abs(f(d - 1e-9) - f(d + 1e-9)) < 1e-7
    assert f(0.0) == base and f(1e6) == cool


@settings(max_examples=40, deadline=None)
@given(nx=st.integers(3, 400), nt=st.integers(4, 5000), xfunc=st.sampled_from(["identity", "sin"]))
def test_spacetime_grid_properties(nx, nt, xfunc):
    g = ebm.SpaceTime(nx, nt, 2, xfunc)
    assert len(g.x) == nx and len(g.t) == nt and np.all(np.diff(g.x) > 0) and 0 < g.x[0] and g.x[-1] < 1
    assert abs(g.t[0] - 0.5 / nt) < 1e-15 and abs(g.t[-1] - (1 - 0.5 / nt)) < 1e-15
    assert 1 <= g.winter.inx <= g.summer.inx <= nt
    # round-half-to-even like Julia's round(Int, x) (infrastructure.jl:131-132)
    v = nt * 0.26125
    if abs(v - math.floor(v) - 0.5) < 1e-9:
        assert g.winter.inx % 2 == 0
    assert abs(g.T(1) - g.t[0]) < 1e-15 and abs(g.T(nt + 1) - (1 + g.t[0])) < 1e-12


@settings(max_examples=40, deadline=None)
@given(nx=st.integers(3, 200), seed=st.integers(0, 10_000))
def test_hemispheric_mean_matches_oracle_and_is_linear(nx, seed):
    rng = np.random.default_rng(seed)
    x = ebm.SpaceTime(nx, 10, 1, "sin").x
    a, b = rng.normal(size=nx), rng.normal(size=nx)
    ha, hb = ebm.hemispheric_mean(a, x), ebm.hemispheric_mean(b, x)
    assert abs(ha - oracle.hemispheric_mean(a, x)) <= 1e-13 * max(1.0, abs(ha))
    assert abs(ebm.hemispheric_mean(2.0 * a - 3.0 * b, x) - (2.0 * ha - 3.0 * hb)) < 1e-12
    assert abs(ebm.hemispheric_mean(np.ones(nx), x) - (x[-1] - x[0])) < 1e-13       # no end caps (utilities.jl:397-403)


@settings(max_examples=80, deadline=None)
@given(total=st.integers(0, 2_000_000), world=st.integers(1, 16))
def test_member_blocks_tile_the_ensemble(total, world):
    blocks = [ebm.member_block(total, world, r) for r in range(world)]
    assert blocks[0][0] == 0 and blocks[-1][0] + blocks[-1][1] == total
    assert all(o1 + c1 == o2 for (o1, c1), (o2, _) in zip(blocks, blocks[1:]))
    assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
