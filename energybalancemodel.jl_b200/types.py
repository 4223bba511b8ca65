"""Host-side mirror of the reference's model / parameter / forcing / solution types.

The reference is Julia; these classes keep the same names, argument meaning and error
behaviour as ``src/infrastructure.jl`` so that the parity tests read like the reference's
own (``test/runtests.jl``) and so that the Julia package extension (``julia/ext``) and this
module marshal exactly the same quantities to the C ABI (``include/ebm_cuda.h``).

Citations (file:line) are relative to the reference repository.
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np

__all__ = [
    "Collection", "SpaceTime", "Forcing", "Solutions", "default_parval", "miz_paramset",
    "classic_paramset", "default_parameters", "CLASSIC_PAR_ORDER", "MIZ_PAR_ORDER",
    "CLASSIC_VARS", "MIZ_VARS", "hemispheric_mean", "annual_mean_forcing", "hysteresis_points",
]


class Collection(dict):
    """``Collection{V}``: a dict with dot access (src/infrastructure.jl:39-49)."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as exc:  # Julia throws KeyError as well
            raise AttributeError(key) from exc

    def __setattr__(self, key, val):
        self[key] = val

    def copy(self):
        return Collection({k: (v.copy() if hasattr(v, "copy") else v) for k, v in self.items()})


def _round_half_even(v: float) -> int:
    # Julia round(Int, x) is round-half-to-even, as is Python's round()
    return int(round(v))


class SpaceTime:
    """``SpaceTime{F}(nx, nt, dur)`` (src/infrastructure.jl:109-141).

    ``xfunc`` is ``"identity"`` (default, ``urange=(0, 1)``) or ``"sin"`` (``urange=(0, pi/2)``).
    Ranges in Julia are TwicePrecision: ``u``, ``t`` and ``T`` are the correctly rounded
    rationals, never ``i*dt`` accumulated in Float64.
    """

    def __init__(self, nx: int, nt: int, dur: int, xfunc: str = "identity", urange=None,
                 winter: float = 0.26125, summer: float = 0.77375):
        if xfunc not in ("identity", "sin"):
            raise ValueError(f"xfunc must be 'identity' or 'sin', got {xfunc!r}")
        if urange is None:
            urange = (0.0, 1.0) if xfunc == "identity" else (0.0, math.pi / 2.0)
        self.xfunc = xfunc
        self.nx, self.nt, self.dur = int(nx), int(nt), int(dur)
        dx = (urange[1] - urange[0]) / nx                                   # :125
        if tuple(urange) == (0.0, 1.0):
            u = [(2 * j - 1) / (2 * nx) for j in range(1, nx + 1)]          # rationalised range (:126)
        else:
            a, st = Fraction(urange[0] + dx / 2.0), Fraction(dx)            # Float64-exact TwicePrecision range
            u = [float(a + j * st) for j in range(nx)]
        self.u = np.array(u, dtype=np.float64)
        self.x = self.u.copy() if xfunc == "identity" else np.array([math.sin(v) for v in u])  # :127
        self.dt = 1.0 / nt                                                  # :128
        self.t = np.array([(2 * i - 1) / (2 * nt) for i in range(1, nt + 1)], dtype=np.float64)  # :129
        self.winter = Collection(t=winter, inx=_round_half_even(nt * winter))   # :131
        self.summer = Collection(t=summer, inx=_round_half_even(nt * summer))   # :132

    def T(self, tinx: int) -> float:
        """``st.T[tinx]`` (1-based): dt/2 : dt : dur - dt/2 (:130)."""
        return (2 * tinx - 1) / (2 * self.nt)

    @property
    def grid_kind(self) -> int:
        # classic/MIZ on SpaceTime{identity} use get_diffop; everything else the generic stencil
        return 0 if (self.xfunc == "identity") else 1

    def __repr__(self):
        return f"SpaceTime{{{self.xfunc}}}({self.nx}, {self.nt}, {self.dur})"


class Forcing:
    """``Forcing(base)`` or ``Forcing(base, peak, cool, holdyrs, rates)`` (src/infrastructure.jl:208-241)."""

    def __init__(self, base: float, peak: float | None = None, cool: float | None = None,
                 holdyrs=(0, 0), rates=(0.0, 0.0)):
        base = float(base)
        if peak is None:
            self.constant = True
            self.base = self.peak = self.cool = base
            self.holdyrs, self.rates, self.domain = (0, 0), (0.0, 0.0), (0, 0, 0, 0, 0)
            return
        self.constant = False
        peak, cool = float(peak), float(cool)
        dom = [0, 0, 0, 0, 0]
        for i in range(1, 5):
            dom[i] += holdyrs[0]
        warming = (peak - base) / rates[0] if rates[0] != 0 else math.inf
        if not (rates[0] > 0 and float(warming).is_integer()):
            raise ValueError(f"Warming time must be positive integer. Got {warming} y.")   # ArgumentError :231
        for i in range(2, 5):
            dom[i] += int(warming)
        for i in range(3, 5):
            dom[i] += holdyrs[1]
        cooling = (cool - peak) / rates[1] if rates[1] != 0 else math.inf
        if not (rates[1] < 0 and float(cooling).is_integer()):
            raise ValueError(f"Cooling time must be positive integer. Got {cooling} y.")   # ArgumentError :238
        dom[4] += int(cooling)
        self.base, self.peak, self.cool = base, peak, cool
        self.holdyrs, self.rates, self.domain = tuple(holdyrs), tuple(map(float, rates)), tuple(dom)

    def __call__(self, T: float) -> float:                                   # :294-307
        if self.constant:
            return self.base
        d = self.domain
        if T < d[1]:
            return self.base
        elif T < d[2]:
            return self.base + self.rates[0] * (T - d[1])
        elif T < d[3]:
            return self.peak
        elif T < d[4]:
            return self.peak + self.rates[1] * (T - d[3])
        return self.cool

    def row(self) -> np.ndarray:
        """Row for the C ABI: base, peak, cool, rate_up, rate_down, 5 breakpoints."""
        return np.array([self.base, self.peak, self.cool, self.rates[0], self.rates[1], *map(float, self.domain)])

    def __repr__(self):
        return f"Forcing({self.base})" if self.constant else f"Forcing({self.base} ↗ {self.peak} ↘ {self.cool})"


# src/infrastructure.jl:407-433
default_parval = Collection(
    D=0.6, A=193.0, B=2.1, cw=9.8, S0=420.0, S1=338.0, S2=240.0, a0=0.7, a2=0.1, ai=0.4, Fb=4.0,
    k=2.0, Lf=9.5, F=0.0, cg=0.01 * 9.8, tau=1e-5, Tm=0.0, m1=1.6e-6 * 31536000, m2=1.36,
    alpha=0.66, rl=0.5, Dmin=1.0, Dmax=156.0, hmin=0.1, kappa=0.01 * 31536000,
)
# :436-444
miz_paramset = ("D", "A", "B", "cw", "S0", "S1", "S2", "a0", "a2", "ai", "Fb", "k", "Lf", "Tm", "m1", "m2",
                "alpha", "rl", "Dmin", "Dmax", "hmin", "kappa")
classic_paramset = ("D", "A", "B", "cw", "S0", "S1", "S2", "a0", "a2", "ai", "Fb", "k", "Lf", "F", "cg", "tau")

# order of the parameter rows crossing the C ABI (ebm_classic_params_t / ebm_miz_params_t)
CLASSIC_PAR_ORDER = ("D", "A", "B", "cw", "S0", "S1", "S2", "a0", "a2", "ai", "Fb", "k", "Lf", "cg", "tau")
MIZ_PAR_ORDER = miz_paramset
# stored variables (src/infrastructure.jl:621-624; order of src/EnergyBalanceModel.jl:63 for MIZ)
CLASSIC_VARS = ("E", "T", "h")
MIZ_VARS = ("T", "Ei", "Ti", "D", "n", "h", "phi", "E", "Ew", "Tw")


def default_parameters(model) -> Collection:
    """``default_parameters(:MIZ)`` / ``default_parameters(:Classic)`` (src/infrastructure.jl:447-474).

    Like the reference, anything that is not ``"MIZ"`` selects the classic set.
    """
    names = miz_paramset if str(model).lstrip(":") == "MIZ" else classic_paramset
    return Collection({k: float(default_parval[k]) for k in names})


def hemispheric_mean(vec, x) -> float:
    """src/utilities.jl:397-403 (trapezoid over cell centres, no end caps)."""
    acc = 0.0
    for i in range(len(x) - 1):
        acc += (vec[i] + vec[i + 1]) * (x[i + 1] - x[i]) / 2.0
    return acc


def annual_mean_forcing(forcing: "Forcing", st: "SpaceTime", year: int) -> float:
    """``annual_mean(forcing, st, year)`` (src/infrastructure.jl:546-547): mean of ``forcing.(year-1 .+ st.t)``."""
    return float(np.mean([forcing((year - 1) + float(t)) for t in st.t]))


def hysteresis_points(diag, season: int = 2):
    """The points ``plot_seasonal`` draws (src/plot.jl:173-190) from the L0 diagnostics of an ensemble run:
    x = hemispheric mean of the annual-mean temperature, y = ice-covered area ``2*pi*hemispheric_mean(phi)`` (MIZ) or
    ``2*pi*hemispheric_mean(E < 0)`` (classic) of ``season`` (0 winter, 1 summer, 2 annual mean).
    ``diag`` is ``[nmem, dur, 3, 4]``; returns two ``[nmem, dur]`` arrays."""
    diag = np.asarray(diag)
    return diag[:, :, 2, 0], diag[:, :, season, 2]


class Solutions:
    """``Solutions{F,C}`` (src/infrastructure.jl:333-383).

    ``raw[var]`` is an array ``[len(ts), nx]`` (``raw.E[ti]`` in Julia is row ``ti-1`` here) and
    ``seasonal.winter/summer/avg[var]`` is ``[dur, nx]``; entries the reference never assigns
    (``undef``) are NaN.
    """

    def __init__(self, st: SpaceTime, forcing: Forcing, par: Collection, init: Collection, variables,
                 lastonly: bool = True):
        self.spacetime, self.forcing, self.parameters, self.initconds = st, forcing, par, init
        self.lastonly, self.debug = lastonly, None
        nt, dur = st.nt, st.dur
        if lastonly:
            self.ts = np.array([(dur - 1) + (2 * i - 1) / (2 * nt) for i in range(1, nt + 1)])   # :353
        else:
            self.ts = np.array([(2 * i - 1) / (2 * nt) for i in range(1, nt * dur + 1)])          # :356
        self.raw = Collection({v: np.full((len(self.ts), st.nx), np.nan) for v in variables})
        self.seasonal = Collection(
            winter=Collection({v: np.full((dur, st.nx), np.nan) for v in variables}),
            summer=Collection({v: np.full((dur, st.nx), np.nan) for v in variables}),
            avg=Collection({v: np.full((dur, st.nx), np.nan) for v in variables}),
        )

    def __repr__(self):
        st = self.spacetime
        return (f"Solutions{{{st.xfunc}, {str(self.forcing.constant).lower()}}}({st.nx}×{len(self.ts)}"
                f"@({self.ts[0]}:{st.dt}:{self.ts[-1]}), {sorted(self.raw)})")
