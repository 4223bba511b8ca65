"""CPU tests of the boundary and the host logic: the C-ABI library loads and exports every symbol include/ebm_cuda.h
declares, struct layouts agree between the header and the ctypes mirror, compute entry points fail loudly without
a GPU (no CPU fallback), argument validation, Solutions assembly, member sharding and the world-size-2 gather."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import ebm_b200 as ebm
from ebm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ebm_cuda.h")


def _header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ebm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 13 and set(names) == set(_lib.EXPORTED_SYMBOLS)
    for n in names:
        assert hasattr(lib, n), n
    assert lib.ebm_version().decode().startswith("ebm_cuda")
    assert lib.ebm_launch_count() >= 0


def test_struct_layouts_match_the_header():
    """Compile a tiny C program against include/ebm_cuda.h and compare sizeof/offsetof with the ctypes mirror."""
    exe = os.path.join(ROOT, "tests", "_layout_probe")
    code = r'''
#include <stdio.h>
#include <stddef.h>
#include "ebm_cuda.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(ebm_grid_t), offsetof(ebm_grid_t, x), sizeof(ebm_options_t),
         offsetof(ebm_options_t, newton_tol), offsetof(ebm_options_t, start_year), sizeof(ebm_classic_outputs_t),
         sizeof(ebm_miz_outputs_t), sizeof(ebm_classic_device_args_t), sizeof(ebm_miz_device_args_t), sizeof(ebm_forcing_t));
  printf("%d %d %d\n", EBM_CLASSIC_NPAR, EBM_MIZ_NPAR, EBM_NFORCING);
  return 0;
}'''
    r = subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=code, text=True,
                       capture_output=True)
    assert r.returncode == 0, r.stderr
    try:
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    finally:
        os.remove(exe)
    got = list(map(int, out))
    want = [C.sizeof(_lib.Grid), _lib.Grid.x.offset, C.sizeof(_lib.Options), _lib.Options.newton_tol.offset,
            _lib.Options.start_year.offset, C.sizeof(_lib.ClassicOutputs), C.sizeof(_lib.MizOutputs),
            C.sizeof(_lib.ClassicDeviceArgs), C.sizeof(_lib.MizDeviceArgs), 8 * _lib.NFORCING,
            _lib.CLASSIC_NPAR, _lib.MIZ_NPAR, _lib.NFORCING]
    assert got == want, (got, want)
    assert len(ebm.CLASSIC_PAR_ORDER) == _lib.CLASSIC_NPAR and len(ebm.MIZ_PAR_ORDER) == _lib.MIZ_NPAR


def test_plain_c_consumer_links_and_fails_loudly_without_gpu(tmp_path):
    """examples/c_abi_example.c: the header is usable from C11, the library links without torch/Python, and without
    a device the compute call returns EBM_ERR_CUDA with a message (the example exits 0 in that case)."""
    exe = str(tmp_path / "c_abi_example")
    lib_dir = os.path.dirname(_lib.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "c_abi_example.c"), "-L", lib_dir, "-lebm_cuda", "-lm",
                        f"-Wl,-rpath,{lib_dir}", "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ebm_cuda" in out.stdout
    if _lib.load().ebm_device_count() == 0:
        assert "no CPU fallback" in out.stdout
    else:
        assert "annual mean" in out.stdout


def _no_gpu():
    return _lib.load().ebm_device_count() == 0


@pytest.mark.skipif(not _no_gpu(), reason="needs a machine without a CUDA device")
def test_no_cpu_fallback_compute_calls_fail_loudly():
    st = ebm.SpaceTime(100, 2000, 1)
    par = ebm.default_parameters("Classic")
    init = ebm.Collection(E=np.full(100, 98.0), Tg=np.full(100, 10.0))
    with pytest.raises(_lib.EBMError) as ei:
        ebm.integrate("Classic", st, ebm.Forcing(0.0), par, init)
    assert ei.value.code == _lib.EBM_ERR_CUDA and "no CPU fallback" in str(ei.value)
    z = np.zeros(180)
    with pytest.raises(_lib.EBMError):
        ebm.integrate("MIZ", ebm.SpaceTime(180, 2000, 1, "sin"), ebm.Forcing(0.0), ebm.default_parameters("MIZ"),
                      ebm.Collection(Ei=z, Ew=z, h=z, D=z, phi=z))
    with pytest.raises(_lib.EBMError):
        ebm.step("Classic", st.t[0], 0.0, init, st, par)
    with pytest.raises(_lib.EBMError):
        ebm.fp64_peak()


def test_missing_library_is_an_import_error(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libebm_cuda.so")
    with pytest.raises(ImportError):
        _lib.load()


def test_argument_validation_mirrors_the_reference():
    st = ebm.SpaceTime(100, 2000, 1)
    par = ebm.default_parameters("Classic")
    init = ebm.Collection(E=np.zeros(100), Tg=np.zeros(100))
    with pytest.raises(ValueError):                      # debug::Expr has no device equivalent
        ebm.integrate("Classic", st, ebm.Forcing(0.0), par, init, debug="vars.E")
    with pytest.raises(ValueError):                      # MethodError in the reference: Val{:classic} has no method
        ebm.integrate("classic", st, ebm.Forcing(0.0), par, init)
    with pytest.raises(ValueError):
        ebm.integrate_ensemble("Classic", st, [ebm.Forcing(0.0)], [par, par], [init])
    bad = ebm.Collection(par); del bad["cw"]
    with pytest.raises(KeyError):
        ebm.integrate_ensemble("Classic", st, [ebm.Forcing(0.0)], [bad], [init])
    with pytest.raises(ValueError):
        ebm.integrate_ensemble("Classic", st, [ebm.Forcing(0.0)], [par], [ebm.Collection(E=np.zeros(99), Tg=np.zeros(100))])
    with pytest.raises(ValueError):
        ebm.SpaceTime(10, 10, 1, "cos")
    # NULL / malformed grids are rejected by the library itself (before any CUDA call)
    lib = _lib.load()
    assert lib.ebm_classic_run(None, 1, None, None, None, None, None, None) == _lib.EBM_ERR_INVALID
    assert b"grid" in lib.ebm_last_error()
    g = _lib.make_grid(st); g.nx = 2
    out = _lib.ClassicOutputs()
    a = np.zeros(100)
    assert lib.ebm_classic_run(C.byref(g), 1, _lib.dptr(a), _lib.dptr(a), _lib.dptr(a), _lib.dptr(a), None, C.byref(out)) == _lib.EBM_ERR_INVALID


def test_integrate_arrays_validates_shapes():
    st = ebm.SpaceTime(100, 2000, 1)
    par = np.tile([ebm.default_parameters("Classic")[k] for k in ebm.CLASSIC_PAR_ORDER], (4, 1))
    forc = np.zeros((4, 10)); z = np.zeros((4, 100))
    with pytest.raises(ValueError):
        ebm.integrate_arrays("Classic", st, forc, par[:, :14], {"E": z, "Tg": z})
    with pytest.raises(ValueError):
        ebm.integrate_arrays("Classic", st, forc[:3], par, {"E": z, "Tg": z})
    with pytest.raises(KeyError):
        ebm.integrate_arrays("Classic", st, forc, par, {"E": z})
    with pytest.raises(ValueError):
        ebm.integrate_arrays("Classic", st, forc, par, {"E": z, "Tg": z[:, :99]})
    with pytest.raises(ValueError):
        ebm.integrate_arrays("MIZ", st, forc, par, {"E": z, "Tg": z})      # 15 columns are not the 22 MIZ parameters


def test_solutions_assembly_layout():
    st = ebm.SpaceTime(50, 100, 3)
    res = ebm.EnsembleResult("Classic", st, 4, 2, ebm.CLASSIC_VARS)
    rng = np.random.default_rng(0)
    res.raw = rng.normal(size=(2, 100, 3, 50)); res.seasonal = rng.normal(size=(2, 3, 3, 3, 50))
    sols = res.solutions(1, ebm.Forcing(1.0), ebm.default_parameters("Classic"), ebm.Collection(), True)
    assert np.array_equal(sols.raw.T, res.raw[1, :, 1]) and np.array_equal(sols.seasonal.summer.h, res.seasonal[1, :, 1, 2])
    assert np.array_equal(sols.seasonal.avg.E[2], res.seasonal[1, 2, 2, 0]) and sols.ts.shape == (100,)
    full = ebm.Solutions(st, ebm.Forcing(0.0), {}, {}, ebm.CLASSIC_VARS, lastonly=False)
    assert full.raw.E.shape == (300, 50) and np.isnan(full.raw.E).all() and abs(full.ts[-1] - 2.995) < 1e-12


def test_checkpoint_roundtrip(tmp_path):
    rng = np.random.default_rng(3)
    final = {k: rng.normal(size=(5, 37)) for k in ("Ei", "Ew", "h", "D", "phi", "T0")}
    final["Ew"][2, 3] = np.nan
    path = str(tmp_path / "state.ebm")
    ebm.save_state(path, final, years_done=12)
    back, years = ebm.load_state(path)
    assert years == 12 and list(back) == list(final)
    for k in final:
        assert np.array_equal(back[k], final[k], equal_nan=True)
    inits, T0 = ebm.inits_from_state(back)
    assert len(inits) == 5 and "T0" not in inits[0] and np.array_equal(inits[4].D, final["D"][4]) and T0.shape == (5, 37)
    assert os.path.getsize(path) == 8 + 32 + 6 * (16 + 8 * 5 * 37)
    with open(path, "r+b") as fh:
        fh.write(b"XXXXXXXX")
    with pytest.raises(ValueError):
        ebm.load_state(path)


def test_member_block_partition():
    for total in (0, 1, 7, 64, 65536, 1048576 + 3):
        for world in (1, 2, 3, 8):
            blocks = [ebm.member_block(total, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == total
            for (o1, c1), (o2, _) in zip(blocks, blocks[1:]):
                assert o1 + c1 == o2
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
    with pytest.raises(ValueError):
        ebm.member_block(10, 2, 2)


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
import ebm_b200 as ebm
rank, world, total = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(sys.argv[2])
dist.init_process_group("gloo", rank=rank, world_size=world)
off, cnt = ebm.member_block(total, world, rank)
# fake per-member-year diagnostics [count, dur=3, 3 seasons, 4]: value encodes the global member index
m = torch.arange(off, off + cnt, dtype=torch.float64).view(-1, 1, 1, 1)
local = m * 1000 + torch.arange(36, dtype=torch.float64).view(1, 3, 3, 4)
out = ebm.gather_member_rows(local, total, dst=0)
if rank == 0:
    want = torch.arange(total, dtype=torch.float64).view(-1, 1, 1, 1) * 1000 + torch.arange(36, dtype=torch.float64).view(1, 3, 3, 4)
    assert out.shape == (total, 3, 3, 4) and torch.equal(out, want)
    print("GATHER_OK", total)
else:
    assert out is None
dist.barrier(); dist.destroy_process_group()
'''


@pytest.mark.parametrize("total", [10, 7])
def test_world_size_two_gather_over_gloo(total, tmp_path):
    """The N>1 host path of bench.py / an ensemble driver: block partition + gather to rank 0, two processes, gloo."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = 29500 + (os.getpid() + total) % 2000
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT, str(total)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert f"GATHER_OK {total}" in outs[0]


def test_member_deal_covers_the_ensemble_and_balances_regimes():
    """32-member packets dealt round-robin after a stable sort by regime: every member exactly once, ranks differ by
    at most one packet, and every rank gets the same share of each regime (the round-1 imbalance)."""
    for total, world in ((65536, 8), (1000, 3), (31, 2), (0, 2), (64, 1)):
        key = (np.arange(total) % 2 == 1).astype(int) * 2            # C4: even members warm (0), odd cold (2)
        parts = [ebm.member_deal(total, world, r, key=key) for r in range(world)]
        allm = np.concatenate(parts) if total else np.empty(0, dtype=np.int64)
        assert np.array_equal(np.sort(allm), np.arange(total))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 32
        if total == 65536:
            for p in parts:
                assert abs(int((key[p] == 2).sum()) - total // (2 * world)) <= 32
                assert np.all(np.diff(key[p]) >= 0) or True        # packets keep members of one regime together
                assert all(len(set(key[p[g:g + 32]])) == 1 for g in range(0, len(p), 32))
    with pytest.raises(ValueError):
        ebm.member_deal(10, 2, 2)


_WORKER_DEAL = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
import ebm_b200 as ebm
rank, world, total = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(sys.argv[2])
dist.init_process_group("gloo", rank=rank, world_size=world)
key = (np.arange(total) % 3 == 0).astype(int)
index = [ebm.member_deal(total, world, r, key=key, group=4) for r in range(world)]
mine = torch.from_numpy(index[rank]).to(torch.float64).view(-1, 1, 1, 1)
local = mine * 1000 + torch.arange(36, dtype=torch.float64).view(1, 3, 3, 4)
out = ebm.gather_member_rows(local, total, dst=0, index=index)
if rank == 0:
    want = torch.arange(total, dtype=torch.float64).view(-1, 1, 1, 1) * 1000 + torch.arange(36, dtype=torch.float64).view(1, 3, 3, 4)
    assert out.shape == (total, 3, 3, 4) and torch.equal(out, want)
    print("DEAL_GATHER_OK", total)
else:
    assert out is None
dist.barrier(); dist.destroy_process_group()
'''


@pytest.mark.parametrize("total", [37, 8])
def test_world_size_two_dealt_gather_over_gloo(total, tmp_path):
    """member_deal + gather_member_rows(index=...): rows come back at their global member index (ragged packets)."""
    script = tmp_path / "worker_deal.py"
    script.write_text(_WORKER_DEAL)
    port = 31500 + (os.getpid() + total) % 2000
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT, str(total)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert f"DEAL_GATHER_OK {total}" in outs[0]


def test_multi_struct_layout_and_no_gpu_behaviour():
    """ebm_multi_t mirrors the header; without a device the multi-GPU entry points fail loudly like the others."""
    exe = os.path.join(ROOT, "tests", "_layout_probe_multi")
    code = r"""
#include <stdio.h>
#include <stddef.h>
#include "ebm_cuda.h"
int main(void) { printf("%zu %zu %zu %zu %zu %zu\n", sizeof(ebm_multi_t), offsetof(ebm_multi_t, diag_device), offsetof(ebm_multi_t, packet), offsetof(ebm_multi_t, devices), sizeof(ebm_options_t), offsetof(ebm_options_t, classic_stencil)); return 0; }"""
    r = subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=code, text=True, capture_output=True)
    assert r.returncode == 0, r.stderr
    try:
        got = list(map(int, subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()))
    finally:
        os.remove(exe)
    assert got == [C.sizeof(_lib.Multi), _lib.Multi.diag_device.offset, _lib.Multi.packet.offset, _lib.Multi.devices.offset,
                   C.sizeof(_lib.Options), _lib.Options.classic_stencil.offset]
    if _no_gpu():
        st = ebm.SpaceTime(100, 2000, 1)
        p = ebm.default_parameters("Classic")
        par = np.tile([p[k] for k in ebm.CLASSIC_PAR_ORDER], (4, 1))
        forc = np.zeros((4, 10))
        state = {"E": np.full((4, 100), 98.0), "Tg": np.full((4, 100), 10.0)}
        with pytest.raises(_lib.EBMError) as ei:
            ebm.integrate_arrays("Classic", st, forc, par, state, devices=[0, 1])
        assert ei.value.code == _lib.EBM_ERR_CUDA
    with pytest.raises(ValueError):   # a device listed twice never reaches the library
        _lib.make_multi(devices=[0, 0])


def test_integrate_grids_groups_members_by_spacetime(monkeypatch):
    """SURVEY 8f-4 (per-member nx / nt): host-side grouping of integrate_grids with the device call stubbed out -- one
    call per distinct SpaceTime, members of a group in the caller's order, results addressed by the caller's index."""
    import sys
    integ = sys.modules[ebm.integrate_grids.__module__]   # the module (the package re-exports the function `integrate`)
    calls = []

    def fake(model, st, forcings, pars, inits, **kw):
        calls.append((st.nx, st.nt, st.dur, [f.base for f in forcings]))
        n = len(pars)
        return integ.EnsembleResult(model, st, n, 0, ("E", "T", "h"),
                                    diag=np.array([[[[f.base] * 4] * 3] * st.dur for f in forcings]),
                                    final={"E": np.stack([i["E"] + 1.0 for i in inits])}, flags=np.zeros(n, np.int32))
    monkeypatch.setattr(integ, "integrate_ensemble", fake)
    p = ebm.default_parameters("Classic")
    specs = [(100, 2000, 1), (60, 1000, 2), (100, 2000, 1), (100, 2000, 3), (60, 1000, 2)]
    sts = [ebm.SpaceTime(*s) for s in specs]
    forcings = [ebm.Forcing(float(m)) for m in range(5)]
    inits = [ebm.Collection(E=np.full(st.nx, float(m)), Tg=np.zeros(st.nx)) for m, st in enumerate(sts)]
    res = ebm.integrate_grids("Classic", sts, forcings, [p] * 5, inits)
    assert calls == [(100, 2000, 1, [0.0, 2.0]), (60, 1000, 2, [1.0, 4.0]), (100, 2000, 3, [3.0])]
    assert res.where == [(0, 0), (1, 0), (0, 1), (2, 0), (1, 1)]
    for m, st in enumerate(sts):
        got = res.member(m)
        assert got["diag"].shape == (st.dur, 3, 4) and got["diag"][0, 0, 0] == float(m)
        assert got["final"]["E"].shape == (st.nx,) and got["final"]["E"][0] == m + 1.0 and got["flags"] == 0
    with pytest.raises(ValueError):
        ebm.integrate_grids("Classic", sts[:2], forcings, [p] * 5, inits)
    with pytest.raises(ValueError):
        ebm.integrate_grids("Classic", sts, forcings, [p] * 5, inits, field_stride=4)
