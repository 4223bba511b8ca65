"""Restart files for ensemble runs.

The reference restarts by passing the last stored state as ``init`` (``save``/``load!`` of a ``Solutions``,
src/io.jl:37-52, 84-92) -- which cannot restart the classic model exactly, because ``Tg`` is not a stored variable
(src/infrastructure.jl:621), nor carry the MIZ closure's warm start (src/miz.jl:47,64).  ``EnsembleResult.final``
holds the complete state (classic ``E, Tg``; MIZ ``Ei, Ew, h, D, phi, T0``); this module writes it to a flat
little-endian file that Julia reads without any package::

    magic "EBMCKPT1" | int64 nmem | int64 nx | int64 nvar | int64 years_done | nvar x (16-byte name, nmem*nx float64)

Julia: ``open(path) do io; read(io, 8); nmem, nx, nvar, years = ntuple(_ -> read(io, Int64), 4);
[(strip(String(read(io, 16)), '\\0'), permutedims(reshape(reinterpret(Float64, read(io, 8nmem*nx)), nx, nmem))) for _ in 1:nvar]; end``.
"""
from __future__ import annotations

import struct

import numpy as np

from .types import Collection

MAGIC = b"EBMCKPT1"
__all__ = ["save_state", "load_state", "inits_from_state"]


def save_state(path: str, final: dict, years_done: int = 0) -> None:
    """Write ``EnsembleResult.final`` (dict of ``[nmem, nx]`` arrays)."""
    names = list(final)
    if not names:
        raise ValueError("empty state")
    nmem, nx = np.asarray(final[names[0]]).shape
    with open(path, "wb") as fh:
        fh.write(MAGIC)
        fh.write(struct.pack("<4q", nmem, nx, len(names), int(years_done)))
        for k in names:
            a = np.ascontiguousarray(final[k], dtype="<f8")
            if a.shape != (nmem, nx):
                raise ValueError(f"state variable {k!r} has shape {a.shape}, expected {(nmem, nx)}")
            fh.write(k.encode("ascii")[:16].ljust(16, b"\0"))
            fh.write(a.tobytes())


def load_state(path: str):
    """Returns ``(final, years_done)``."""
    with open(path, "rb") as fh:
        if fh.read(8) != MAGIC:
            raise ValueError(f"{path}: not an EBM checkpoint")
        nmem, nx, nvar, years = struct.unpack("<4q", fh.read(32))
        final = {}
        for _ in range(nvar):
            name = fh.read(16).rstrip(b"\0").decode("ascii")
            buf = fh.read(8 * nmem * nx)
            if len(buf) != 8 * nmem * nx:
                raise ValueError(f"{path}: truncated")
            final[name] = np.frombuffer(buf, dtype="<f8").reshape(nmem, nx).copy()
    return final, years


def inits_from_state(final: dict):
    """``(inits, T0guess)`` to pass back to ``integrate_ensemble`` (``T0guess`` is None for the classic model)."""
    keys = [k for k in final if k != "T0"]
    nmem = next(iter(final.values())).shape[0]
    inits = [Collection({k: final[k][m].copy() for k in keys}) for m in range(nmem)]
    return inits, final.get("T0")
