"""The reference's OWN CUDA kernels (Splines/gpuSplineUtils.cu, built from /root/reference into
oracle/_ref) run on the B200 next to the oracle and next to libm3b200: pins the spline-weight part
of the path to the reference itself."""
import numpy as np
import pytest

from mach3_b200 import handlers, synth
from oracle import binding as O
from oracle import ref_gpu_binding as R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("wl,n", [("CFG1", 20_000), ("SPARSE", 20_000), ("SPARSE_RUNS", 20_000)])
def test_reference_cuda_kernels_vs_oracle_vs_b200(wl, n):
    w = getattr(synth, wl).scaled(n)
    if not R.available(w.n_params):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    typ, npts, cx = synth.param_layout(w)
    spl = synth.make_splines(w)
    ref = R.RefSMonolithGPU(w.n_params, w.n_knots, cx, spl)
    O.set_multithread(False)
    omono = O.SMonolith(w.n_params, w.n_knots, cx, npts, spl)
    gmono = handlers.SMonolith(w.n_params, w.n_knots, cx, npts, spl)
    pars = np.zeros(w.n_params)
    gmono.setSplinePointers(pars)
    for step in (-1, 0, 1, -2, 2, -3, -4):
        pars[:] = synth.proposal(w, step)[0]
        omono.set_params(pars); omono.Evaluate()
        w_ref = ref.run(omono.param_values, omono.segments)
        gmono.Evaluate(); gmono.SynchroniseMemTransfer()
        # the reference's kernels multiply left to right like its serial CPU build: all three bit-exact
        np.testing.assert_array_equal(omono.total_weights, w_ref)
        np.testing.assert_array_equal(gmono.cpu_total_weights, w_ref)
    O.set_multithread(True)
    ref.close()
