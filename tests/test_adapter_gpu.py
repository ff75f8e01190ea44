"""Drop-in check at the SMonolithGPU seam: adapters/SMonolithGPU_m3b200.cu implements the
reference's class (declared by the reference's OWN header Splines/gpuSplineUtils.cuh) on top of
libm3b200's C ABI.  The same harness that drives the reference's CUDA kernels the way SMonolith does
(InitGPU_* -> CopyToGPU_SplineMonolith -> RunGPU_SplineMonolith -> SynchroniseSplines) drives the
adapter; per-event total weights must be bit-identical to the reference kernels' and the oracle's."""
import os

import numpy as np
import pytest

from mach3_b200 import synth
from oracle import binding as O
from oracle import ref_gpu_binding as R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("wl,n", [("CFG1", 20_011), ("SPARSE", 20_000), ("SPARSE_RUNS", 20_000), ("CFG2", 30_000)])
def test_smonolithgpu_adapter_is_a_drop_in(wl, n):
    if not os.path.exists(R.adapter_path()):
        pytest.skip("oracle/_ref/libm3adapter_gpu.so not built (needs /root/reference headers at build time)")
    w = getattr(synth, wl).scaled(n)
    typ, npts, cx = synth.param_layout(w)
    spl = synth.make_splines(w)
    adp = R.RefSMonolithGPU(w.n_params, w.n_knots, cx, spl, adapter=True)
    ref = R.RefSMonolithGPU(w.n_params, w.n_knots, cx, spl) if R.available(w.n_params) else None
    O.set_multithread(False)
    omono = O.SMonolith(w.n_params, w.n_knots, cx, npts, spl)
    pars = np.zeros(w.n_params)
    for step in (-1, 0, 1, -2, 2, -3, -4):
        pars[:] = synth.proposal(w, step)[0]
        omono.set_params(pars); omono.Evaluate()          # FindSplineSegment on the host, like SMonolith::Evaluate
        w_adp = adp.run(omono.param_values, omono.segments)
        np.testing.assert_array_equal(w_adp, omono.total_weights)
        if ref is not None:
            np.testing.assert_array_equal(w_adp, ref.run(omono.param_values, omono.segments))
    O.set_multithread(True)
    adp.close()
    if ref is not None:
        ref.close()


@pytest.mark.parametrize("args", [["30011"], ["20000", "barlow"]])
def test_sample_handler_adapter_runs_the_fitters_call_surface(args):
    """adapters/SampleHandlerB200.h compiled against the mock MaCh3 (tests/adapters/mock_mach3.h): a C++
    program calls Reweight()/GetLikelihood()/GetSampleLikelihood() through SampleHandlerBase pointers on
    a CPU instance and on the B200 adapter wired from the same EventInfo pointer soup."""
    import subprocess
    exe = os.path.join(os.path.dirname(R.adapter_path()), "adapter_test")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/adapter_test not built")
    r = subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "ADAPTER OK" in r.stdout, r.stdout[-3000:]
