"""CPU: host logic of GpuAssembler.assemble_bma_matrices (COO -> CSR, PEC bookkeeping, return values) against the
reference's Assembler.assemble_bma_matrices on a real Electrodynamics3D, with the device element loop replaced by the
oracle restatement (the `-m gpu` drop-in test runs the same comparison with emb_bma_element_matrices).  Needs the
reference package (build container: /root/reference; GPU box: oracle/_ref)."""
import os

import numpy as np
import pytest

from oracle.refharness import harness as H

pytestmark = pytest.mark.skipif(not H.reference_available(), reason="reference package not present")


class _OracleCtx:
    def bma_element_matrices(self, xy, tris, edges, t2e, er, ur, k0):
        from oracle import ned2_oracle as O
        return O.bma_element_matrices(np.asarray(xy), np.asarray(tris), np.asarray(edges), np.asarray(t2e), np.asarray(er),
                                      np.asarray(ur), float(k0))


def test_bma_matrices_and_solve_ids_match_the_reference_assembler():
    from tests.golden import make_golden as G
    from emerge_b200.dropin import GpuAssembler
    box = G.microstrip_box(8, 6, 5)
    fem, phys, mesh, ports = G.modal_physics(box)
    phys.frequencies = [1e9]
    port = ports[1]
    phys.modal_analysis(port, 1, direct=True, TEM=True, freq=1e9)
    k0 = 2 * np.pi * 1e9 / 299792458
    E0, B0, ids0, nlf0 = phys.assembler.assemble_bma_matrices(phys.basis, port._er, port._ur, k0, port, phys.boundary_conditions)
    pece0, pecv0 = list(port._pece), list(port._pecv)
    asm = GpuAssembler.__new__(GpuAssembler)
    asm.ctx = _OracleCtx()
    E1, B1, ids1, nlf1 = asm.assemble_bma_matrices(phys.basis, port._er, port._ur, k0, port, phys.boundary_conditions)
    assert np.array_equal(ids0, ids1) and nlf1.n_field == nlf0.n_field
    assert list(port._pece) == pece0 and list(port._pecv) == pecv0 and port._field is nlf1
    assert abs(E1 - E0).max() <= 1e-12 * abs(E0).max()
    assert abs(B1 - B0).max() <= 1e-12 * abs(B0).max()
