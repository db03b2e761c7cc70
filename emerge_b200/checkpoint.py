"""On-disk checkpoints either side of the hot path (the reference has none - a sweep that dies is lost,
fem/physics/edm/emfreq3d.py:658-732 keeps everything in memory until the end).

  * `SweepCheckpoint`: progress of a frequency sweep - the S-parameters and solver statistics of every point solved so
    far, keyed by a fingerprint of the problem (mesh tables, materials, boundary conditions, frequency list, tolerance).
    `FrequencySweep.run(..., on_point=ckpt.on_point)` appends after every point (atomic rename); a restarted job passes
    `order=ckpt.remaining(order)` and merges `ckpt.S` - solved points are never solved again.
  * `save_operators` / `load_operators`: the assembled E (curl-curl) and B (mass) matrices with their shared pattern,
    as the reference caches them in Assembler.cached_matrices (assembler.py:324-331), for inspection or for handing the
    operators to another tool.  scipy CSR in, scipy CSR out.
"""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np


def problem_fingerprint(tables, er, ur, bcs, freqs, rtol) -> str:
    h = hashlib.sha256()
    for name in ("nodes", "tets", "edges", "tris"):
        a = np.ascontiguousarray(getattr(tables, name))
        h.update(name.encode() + str(a.shape).encode() + a.tobytes())
    for a in (er, ur):
        a = np.ascontiguousarray(a)
        h.update(str(a.shape).encode() + a.tobytes())
    for b in bcs:
        ids = getattr(b, "tri_ids", None)
        tags = getattr(b, "tags", None)
        h.update(type(b).__name__.encode())
        if ids is not None:
            h.update(np.ascontiguousarray(ids).tobytes())
        elif tags is not None:
            h.update(repr(list(np.ravel(tags))).encode())
        h.update(repr(getattr(b, "port_number", None)).encode())
    h.update(np.ascontiguousarray(np.asarray(freqs, dtype=float)).tobytes())
    h.update(repr(float(rtol)).encode())
    return h.hexdigest()


class SweepCheckpoint:
    def __init__(self, path: str, fingerprint: str, freqs, n_ports: int):
        self.path = path
        self.fingerprint = fingerprint
        self.freqs = np.asarray(freqs, dtype=float)
        self.S = np.zeros((len(self.freqs), n_ports, n_ports), dtype=np.complex128)
        self.done = np.zeros(len(self.freqs), dtype=bool)
        self.stats = {}
        if os.path.exists(path):
            self._load()

    def _load(self):
        with np.load(self.path, allow_pickle=False) as z:
            if str(z["fingerprint"]) != self.fingerprint:
                raise ValueError(f"{self.path}: checkpoint belongs to a different problem (fingerprint mismatch)")
            if z["S"].shape != self.S.shape:
                raise ValueError(f"{self.path}: checkpoint has S of shape {z['S'].shape}, expected {self.S.shape}")
            self.S, self.done = z["S"].copy(), z["done"].copy()
            self.stats = {int(k): v for k, v in json.loads(str(z["stats"])).items()}

    def save(self):
        tmp = self.path + ".tmp.npz"
        np.savez(tmp, fingerprint=self.fingerprint, freqs=self.freqs, S=self.S, done=self.done,
                 stats=json.dumps({str(k): v for k, v in self.stats.items()}))
        os.replace(tmp, self.path)          # atomic: a crash leaves the previous checkpoint intact

    def on_point(self, i, S_i, stats):
        """callback for FrequencySweep.run(on_point=...)"""
        self.S[i] = S_i
        self.done[i] = True
        def plain(v):
            if isinstance(v, (bool, np.bool_)):
                return bool(v)
            if isinstance(v, (float, np.floating)):
                return float(v)
            if isinstance(v, (int, np.integer)):
                return int(v)
            return v
        self.stats[int(i)] = [{k: plain(v) for k, v in s.items()
                               if isinstance(v, (int, float, bool, str, np.integer, np.floating, np.bool_))} for s in stats]
        self.save()

    def remaining(self, order) -> list:
        return [int(i) for i in order if not self.done[int(i)]]

    def merge_into(self, S) -> np.ndarray:
        """fills the rows of S (nf, P, P) that this checkpoint holds and returns S"""
        S[self.done] = self.S[self.done]
        return S


def save_operators(path: str, E, B) -> None:
    """E, B: scipy CSR with one shared canonical pattern (optimized_assembly.py:61-62)"""
    E, B = E.tocsr(), B.tocsr()
    if not (np.array_equal(E.indptr, B.indptr) and np.array_equal(E.indices, B.indices)):
        raise ValueError("E and B must share one pattern")
    np.savez(path, shape=np.array(E.shape), indptr=E.indptr.astype(np.int64), indices=E.indices.astype(np.int32),
             E=E.data, B=B.data)


def load_operators(path: str):
    import scipy.sparse as sp
    with np.load(path, allow_pickle=False) as z:
        shape = tuple(int(v) for v in z["shape"])
        E = sp.csr_matrix((z["E"], z["indices"], z["indptr"]), shape=shape)
        B = sp.csr_matrix((z["B"], z["indices"], z["indptr"]), shape=shape)
    return E, B
