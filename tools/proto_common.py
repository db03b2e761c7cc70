"""Scratch helpers for CPU prototyping of the Krylov solver (not product code, not imported by tests)."""
import sys, os, time
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ned2_oracle as O
from emerge_b200.synthmesh import box_mesh, mesh_tables, tri_ids_of

def waveguide_system(nx, ny, nz, L, jitter=0.0, er_fn=None):
    a, b = 22.86e-3, 10.16e-3
    box = box_mesh(nx, ny, nz, a, b, L, jitter=jitter, seed=0)
    t = mesh_tables(box.nodes_xyz, box.tets)
    nT = t.tets.shape[1]
    er = np.repeat(np.eye(3, dtype=complex)[:, :, None], nT, axis=2)
    ur = er.copy()
    if er_fn is not None:
        cen = box.nodes_xyz[box.tets].mean(axis=1)
        er = er * er_fn(cen)[None, None, :]
    E, B = O.assemble_EB(t.nodes, t.tets, t.edges, t.tris, t.edge_lengths, t.tet_to_field, t.tet_to_edge, ur, er)
    N = t.n_field
    ports = []
    for tag, z in ((5, 0.0), (6, L)):
        ids = tri_ids_of(t, box.face_tris[box.face_tag == tag])
        v = t.nodes[:, t.tris[:, ids]]
        x, y = v[0].T, v[1].T
        S = O.gen_csr_tri(N, t.tri_to_field, ids, O.tri_surface_matrix(x, y))
        ports.append(dict(ids=ids, x=x, y=y, S=S))
    pec = []
    for tag in (1, 2, 3, 4):
        ids = tri_ids_of(t, box.face_tris[box.face_tag == tag])
        pec.append(t.tri_to_field[:, ids].ravel())
    pec = np.unique(np.concatenate(pec))
    solve_ids = np.setdiff1d(np.arange(N), pec)
    return dict(box=box, t=t, E=E, B=B, ports=ports, solve_ids=solve_ids, a=a, b=b, L=L, N=N)

def system_at(sys_, f):
    a, b = sys_['a'], sys_['b']
    k0 = 2 * np.pi * f / 299792458
    beta = np.sqrt(k0 ** 2 - (np.pi / a) ** 2 + 0j)
    A = sys_['E'] - sys_['B'] * k0 ** 2
    rhs = []
    DP = O.dunavant4()
    amp = np.sqrt(4 * 377 / (a * b))
    for p in sys_['ports']:
        A = A + (1j * beta) * p['S']
        xq = p['x'] @ DP[1:4]
        U = np.zeros((2, 6, len(p['ids'])), dtype=complex)
        U[1] = (-2j * beta * amp * np.cos(np.pi * xq / a)).T
        bl = O.tri_forcing(p['x'], p['y'], U)
        bv = np.zeros(sys_['N'], dtype=complex)
        np.add.at(bv, sys_['t'].tri_to_field[:, p['ids']].T, bl)
        rhs.append(bv)
    s = sys_['solve_ids']
    As = A.tocsr()[s][:, s].tocsr()
    return As, [r[s] for r in rhs]
