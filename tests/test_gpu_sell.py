"""The staged SELL operator application of the inner iteration (csrc/sell.cuh, k_bsell_tma) checked directly: against the
complex128 operator A(f) it is the symmetric part of (solver.py:405-469 is replaced by an iteration on this operator),
and the tuning variants against each other (every variant adds
the blocks of a row in the same order, so the products are bitwise equal)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cells,slabs", [((8, 4, 12), False), ((12, 6, 24), True), ((24, 12, 80), False)])
def test_sell_product_and_variants(cells, slabs):
    import bench
    from emerge_b200.sweep import FrequencySweep
    box, t, er, ur, bcs, L = bench.make_waveguide(*cells, slabs=slabs)
    sw = FrequencySweep(t, er, ur, bcs, recycle=0)
    sw.solver_opts.update(precond="block")
    sw.setup()
    ctx = sw.ctx
    sw.assemble_frequency(10e9)
    old = {k: os.environ.get(k) for k in ("EMB_SPMV_CHECK", "EMB_SELL_VARIANT")}
    os.environ["EMB_SPMV_CHECK"] = "1"
    try:
        sums = {}
        for v in (None, "8:24:2:0", "16:16:2:1", "5:32:2:1", "8:24:3:1"):
            if v is None:
                os.environ.pop("EMB_SELL_VARIANT", None)
            else:
                os.environ["EMB_SELL_VARIANT"] = v
            ctx.spmv_bench(1, nv=2, fp32=True)
            sums[v] = tuple(ctx.last_ms(k) for k in ("spmv_check_re", "spmv_check_im", "spmv_check_abs2"))
            # As = symmetric part of A(f) in complex64; the reference's mass matrix is asymmetric at the 1e-4 level
            assert 0 <= ctx.last_ms("spmv_check_vs_A") < 2e-3, (v, ctx.last_ms("spmv_check_vs_A"))
        assert len(set(sums.values())) == 1, sums
        assert np.isfinite(sums[None]).all() and sums[None][2] > 0
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        ctx.close()
