"""Stratton-Chu far field on the device: host-side mirror of fem/physics/edm/sc.py.

`stratton_chu(Ein, Hin, mesh, theta, phi, k0)` has the reference's signature and meaning (sc.py:144-180): `mesh` is any
object with the SurfaceMesh attributes the reference reads (`areas`, `edge_centers`, `normals`, `tri_to_edge`, `n_tris`;
fem/mesh3d.py:458-575), Ein / Hin are the fields at `mesh.exyz` (EMDataSet.interpolate(*surf.exyz).EH, as
demo3_patch_antenna.py:96-100 does).  The area-weighted edge normals (sc.py:155-166, a Python loop over the triangles)
are one scatter-add here, the double sum over directions x samples runs in `emb_stratton_chu` (csrc/postproc.cu).
"""
from __future__ import annotations

import numpy as np

from .lib import Context


def weighted_edge_normals(mesh) -> np.ndarray:
    """wns[:, e] = sum over the triangles t that own surface edge e of normal_t * area_t / 3 (sc.py:155-166), float32
    like the reference's accumulator (it is created with .astype(np.float32), sc.py:155)."""
    n_edges = np.asarray(mesh.edge_centers).shape[1]
    w = (np.asarray(mesh.normals, dtype=np.float64) * (np.asarray(mesh.areas, dtype=np.float64) / 3.0)[None, :])
    wns = np.zeros((3, n_edges), dtype=np.float32)
    t2e = np.asarray(mesh.tri_to_edge, dtype=np.int64)
    # the reference adds triangle by triangle, edge 1, 2, 3, in float32: same order here (np.add.at is sequential)
    idx = t2e.T.ravel()                                        # t0e0, t0e1, t0e2, t1e0, ...
    for a in range(3):
        np.add.at(wns[a], idx, np.repeat(w[a], 3).astype(np.float32))
    return wns


def stratton_chu(Ein, Hin, mesh, theta, phi, k0: float, ctx: Context | None = None, device: int = 0):
    """-> (E (3,n), H (3,n)) complex128 far-field patterns at the directions (theta, phi); drop-in for
    fem.physics.edm.stratton_chu (sc.py:144-180).  ctx: an existing library context (e.g. `asm.ctx` of the installed
    assembler); otherwise a temporary one on `device`."""
    Ein = np.asarray(Ein, dtype=np.complex128)
    Hin = np.asarray(Hin, dtype=np.complex128)
    wns = weighted_edge_normals(mesh)
    own = ctx is None
    if own:
        ctx = Context(device)
    try:
        return ctx.stratton_chu(Ein, Hin, np.asarray(mesh.edge_centers, dtype=np.float64), wns, theta, phi, k0)
    finally:
        if own:
            ctx.close()


def install_farfield(edm_module, ctx_getter):
    """Replaces `stratton_chu` of the reference's fem.physics.edm package (and of its sc module) by the device version.
    ctx_getter() -> Context or None."""
    original = getattr(edm_module.stratton_chu, "_reference", edm_module.stratton_chu)

    def patched(Ein, Hin, mesh, theta, phi, k0):
        ctx = ctx_getter()
        if ctx is not None and getattr(ctx, "h", None) is None:        # that assembler's context has been closed
            ctx = None
        return stratton_chu(Ein, Hin, mesh, theta, phi, k0, ctx=ctx)
    patched.__doc__ = stratton_chu.__doc__
    patched._reference = original                                       # the reference's own implementation
    edm_module.stratton_chu = patched
    sc = getattr(edm_module, "sc", None)
    if sc is not None:
        sc.stratton_chu = patched
    return patched
