"""Touchstone (v1.1) export / import of the S-parameters of a sweep.

Data format on the output side of the hot path: the reference's EMSimData.export_touchstone
(fem/physics/edm/emdata.py:306-331) collects Smat[f, i-1, j-1] = S(i, j) over the frequency axis and hands it to
fem/physics/edm/touchstone.py:6-50, which delegates the file format to scikit-rf.  This module writes the same file
without that dependency: option line `# HZ S <RI|MA|DB> R <z0>`, one record per frequency; 2-port files use the
historical column order S11 S21 S12 S22, n-port files (n != 2) are row-major with at most four pairs per line.
"""
from __future__ import annotations

import os

import numpy as np


def _pairs(S, fmt):
    if fmt == "RI":
        return S.real, S.imag
    mag, ang = np.abs(S), np.angle(S, deg=True)
    if fmt == "MA":
        return mag, ang
    if fmt == "DB":
        with np.errstate(divide="ignore"):
            return 20 * np.log10(mag), ang
    raise ValueError(f"data_format must be RI, MA or DB, got {fmt!r}")


def write_touchstone(filename: str, freqs, S, data_format: str = "RI", z0: float = 50.0, comments=()) -> str:
    """freqs (M,) in Hz; S (M, N, N) with S[f, i, j] = S(i+1, j+1) (emdata.py:322-326).  Returns the path written
    ('.sNp' appended when the name has no extension, touchstone.py:41-45)."""
    freqs = np.asarray(freqs, dtype=float)
    S = np.asarray(S, dtype=np.complex128)
    if S.ndim != 3 or S.shape[1] != S.shape[2]:
        raise ValueError(f"S must have shape (M, N, N), got {S.shape}")
    if freqs.ndim != 1 or freqs.size != S.shape[0]:
        raise ValueError(f"freqs must be 1-D of length {S.shape[0]}, got shape {freqs.shape}")
    fmt = data_format.upper()
    n = S.shape[1]
    base, ext = os.path.splitext(filename)
    path = base + (ext if ext else f".s{n}p")
    lines = [f"! {c}" for c in comments]
    lines.append(f"# HZ S {fmt} R {z0:g}")
    for k, f in enumerate(freqs):
        Sk = S[k].T if n == 2 else S[k]           # 2-port: S11 S21 S12 S22
        a, b = _pairs(Sk, fmt)
        vals = [f"{u:.12e} {v:.12e}" for u, v in zip(a.ravel(), b.ravel())]
        if n <= 2:
            lines.append(f"{f:.12e} " + " ".join(vals))
            continue
        for i in range(n):                         # one matrix row at a time, at most four pairs per line
            row = vals[i * n:(i + 1) * n]
            for c0 in range(0, n, 4):
                head = f"{f:.12e} " if (i == 0 and c0 == 0) else " " * 4
                lines.append(head + " ".join(row[c0:c0 + 4]))
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    return path


def read_touchstone(path: str):
    """-> (freqs Hz (M,), S (M, N, N), z0); inverse of write_touchstone (N from the .sNp extension)."""
    ext = os.path.splitext(path)[1].lower()
    n = int(ext[2:-1])
    unit, fmt, z0 = 1e9, "MA", 50.0
    nums = []
    with open(path) as fh:
        for line in fh:
            line = line.split("!")[0].strip()
            if not line:
                continue
            if line.startswith("#"):
                tok = line[1:].upper().split()
                unit = {"HZ": 1.0, "KHZ": 1e3, "MHZ": 1e6, "GHZ": 1e9}[tok[0]]
                fmt = tok[2]
                z0 = float(tok[tok.index("R") + 1]) if "R" in tok else 50.0
                continue
            nums.extend(float(v) for v in line.split())
    rec = 1 + 2 * n * n
    data = np.array(nums).reshape(-1, rec)
    freqs = data[:, 0] * unit
    a, b = data[:, 1::2], data[:, 2::2]
    if fmt == "RI":
        S = a + 1j * b
    elif fmt == "MA":
        S = a * np.exp(1j * np.deg2rad(b))
    else:
        S = 10 ** (a / 20) * np.exp(1j * np.deg2rad(b))
    S = S.reshape(-1, n, n)
    if n == 2:
        S = S.transpose(0, 2, 1)
    return freqs, S, z0


def export_touchstone(result, filename: str, data_format: str = "RI", z0: float = 50.0) -> str:
    """`result`: a SweepResult (emerge_b200.sweep) or a reference EMSimData; ports must be numbered 1..N as the
    reference assumes (emdata.py:311-314)."""
    if hasattr(result, "S") and hasattr(result, "freqs"):
        return write_touchstone(filename, result.freqs, result.S, data_format, z0)
    sets = result.datasets
    n = len(sets[0].Sp.map)
    freqs = np.array([d.freq for d in sets], dtype=float)
    S = np.array([[[d.Sp(i, j) for j in range(1, n + 1)] for i in range(1, n + 1)] for d in sets])
    return write_touchstone(filename, freqs, S, data_format, z0)
