"""Generate tests/golden/*.npz from the C oracle (run here; the fixtures are committed).

The reference's only fixture, test/solution_1year.jld2, is absent from the mount (.MISSING_LARGE_BLOBS) and Julia
is not installed, so these vectors are a STAND-IN produced by the oracle (PARITY UNPINNED, see oracle/ebm_oracle.h):
they pin the oracle against regressions and give the GPU tests something to compare against that does not need
the oracle to be rebuilt.  Layout mirrors the reference's test (test/runtests.jl:37-47): raw[var][10] of the ten MIZ
variables for integrate(:MIZ, SpaceTime{sin}(180,2000,1), Forcing(0.0), default_parameters(:MIZ), zeros).
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ebm_b200 as ebm
from helpers import oracle_classic, oracle_miz, warm_init, cold_init

out = os.path.join(ROOT, "tests", "golden")
os.makedirs(out, exist_ok=True)

# MIZ: the reference fixture's setup; step-10 snapshot + steps 1..20 + end-of-year state
st = ebm.SpaceTime(180, 2000, 1, "sin")
z = np.zeros(180)
init = ebm.Collection(Ei=z, Ew=z, h=z, D=z, phi=z)
o = oracle_miz(st, [ebm.Forcing(0.0)], [ebm.default_parameters("MIZ")], [init], lastonly=False, raw=True)
np.savez_compressed(os.path.join(out, "miz_fixture_setup.npz"),
                    variables=np.array(ebm.MIZ_VARS), step10=o["raw"][0, 9], first20=o["raw"][0, :20],
                    final=np.stack([o[k][0] for k in ("Ei", "Ew", "h", "D", "phi", "T0")]),
                    newton_iters=o["newton_iters"])

# classic: C1a (1 year, warm start) sampled every 100 steps + final state; 30-year seasonal diagnostics fields
st = ebm.SpaceTime(100, 2000, 1)
par = ebm.default_parameters("Classic")
o = oracle_classic(st, [ebm.Forcing(0.0)], [par], [warm_init(100)], lastonly=False, raw=True)
st30 = ebm.SpaceTime(100, 2000, 30)
o30 = oracle_classic(st30, [ebm.Forcing(0.0), ebm.Forcing(0.0)], [par, par], [warm_init(100), cold_init(100)], seasonal=True)
np.savez_compressed(os.path.join(out, "classic_default.npz"),
                    variables=np.array(ebm.CLASSIC_VARS), every100=o["raw"][0, 99::100], final_E=o["E"][0], final_Tg=o["Tg"][0],
                    y30_seasonal=o30["seasonal"][:, 29], y30_E=o30["E"], y30_Tg=o30["Tg"])
print("wrote", os.listdir(out))
