"""B200-native ensemble integrator for the time-stepping path of EnergyBalanceModel.jl.

Importable as ``ebm_b200`` (see the shim package of that name; this directory carries the
reference's name).  The host-side names mirror the reference's exports
(src/EnergyBalanceModel.jl:79-82) for this path: ``Vec``-like NumPy arrays, ``Collection``,
``SpaceTime``, ``Forcing``, ``Solutions``, ``integrate``, ``default_parameters``.  All arithmetic
runs in ``lib/libebm_cuda.so`` (hand-written CUDA for sm_100a) behind the C ABI in
``include/ebm_cuda.h``; without the library or without a GPU every compute call raises.
"""
from .types import (CLASSIC_PAR_ORDER, CLASSIC_VARS, MIZ_PAR_ORDER, MIZ_VARS, Collection, Forcing, Solutions,
                    SpaceTime, annual_mean_forcing, classic_paramset, default_parameters, default_parval,
                    hemispheric_mean, hysteresis_points, miz_paramset)
from .integrate import (EnsembleResult, GroupedResult, fp64_peak, integrate, integrate_arrays, integrate_ensemble,
                        integrate_grids, model_name, step)
from .checkpoint import inits_from_state, load_state, save_state
from .sharding import gather_member_rows, member_block, member_deal
from . import _lib, build  # noqa: F401

__all__ = [
    "Collection", "SpaceTime", "Forcing", "Solutions", "integrate", "integrate_ensemble", "integrate_arrays", "integrate_grids", "GroupedResult", "step",
    "default_parameters", "default_parval", "miz_paramset", "classic_paramset", "hemispheric_mean",
    "annual_mean_forcing", "hysteresis_points", "save_state", "load_state", "inits_from_state",
    "EnsembleResult", "fp64_peak", "member_block", "member_deal", "gather_member_rows", "model_name", "CLASSIC_PAR_ORDER", "MIZ_PAR_ORDER", "CLASSIC_VARS", "MIZ_VARS",
]
