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


@pytest.mark.parametrize("args", [["30011"], ["20000", "barlow"], ["30011", "poisson", "3"], ["9000", "barlow", "8"],
                                  ["20000", "poisson", "1", "batch"], ["8000", "barlow", "1", "batch"]])
def test_sample_handler_adapter_runs_the_fitters_call_surface(args):
    """adapters/SampleHandlerB200.h compiled against the mock MaCh3 (tests/adapters/mock_mach3.h): a C++
    program calls Reweight()/GetLikelihood()/GetSampleLikelihood() through SampleHandlerBase pointers on
    a CPU instance and on the B200 adapter wired from the same EventInfo pointer soup.  A third argument n > 1 spreads
    the sample over n group members (m3b_group_*: one process, one calling thread; devices 0..n-1 modulo the GPUs of
    the box) -- the 8-GPU step from the single-process boundary.  "batch": adapters/BatchFitters.h (RunLLHScan,
    PredictiveThrower's toy loop, DelayedMR2T2's stages on m3b_step_batch) against the sequential loops on the CPU
    instance."""
    import subprocess
    exe = os.path.join(os.path.dirname(R.adapter_path()), "adapter_test")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/adapter_test not built")
    r = subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "ADAPTER OK" in r.stdout, r.stdout[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("n_members", [1, 2])
def test_adapter_over_the_reference_real_sample_handler_class(n_members):
    """adapters/SampleHandlerB200.h instantiated over the reference's REAL SampleHandlerFD (compiled from
    /root/reference, oracle/ref_host) and linked with libm3b200: after MoveToB200() the fitters' calls --
    Reweight(), GetLikelihood(), GetSampleLikelihood() -- run on the B200 and return what the reference's own CPU
    implementation returned for the same object (tests/golden/ref_host_fd.npz), W2 frozen and live."""
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    import refpath_cases as RC
    from oracle import ref_path_binding as RP
    if not RP.available_b200():
        pytest.skip("oracle/_ref/libm3ref_path_lm_b200.so not built (needs /root/reference at build time)")
    gold = np.load(os.path.join(here, "golden", "ref_host_fd.npz"))
    for update_w2 in (False, True):
        tag = f"mono_w2{int(update_w2)}"
        f = RC.fd_case()
        c = f["mono"]
        E = f["sample_id"].size
        m = RP.RefSMonolith(c["type"], c["npts"], c["vals"], build="float_b200")
        fd = RP.RefSampleHandlerFD(RC.fd_edges(), 1, update_w2, build="float_b200")
        fd.attach_monolith(m)
        idx = np.arange(E, dtype=np.int32)
        fd.set_events(f["sample_id"], f["kin"], f["norm_idx"], RC.NPE, RC.N_NORM, w_before=idx, w_after=E + idx, n_pool=2 * E)
        # step 0 on the reference's CPU path (before MoveToB200 the adapter forwards to the base class) ...
        pool = np.concatenate([f["osc"][0], f["static_w"]]).astype(np.float64)
        fd.reweight(f["pars"][0], f["norm"][0], pool)
        np.testing.assert_array_equal(fd.hist()[0], gold[f"{tag}/mc"][0])
        fd.set_data(gold[f"{tag}/data"])
        # ... then the same object moves to the B200 and replays the chain from its first step
        import torch
        fd.move_to_b200(0 if n_members == 1 else [i % torch.cuda.device_count() for i in range(n_members)])
        for t in range(8):                                   # (steps 8+ shift the kinematics: not supported by the adapter)
            pool = np.concatenate([f["osc"][t], f["static_w"]]).astype(np.float64)
            fd.reweight(f["pars"][t], f["norm"][t], pool)
            assert fd.llh() == pytest.approx(float(gold[f"{tag}/llh"][t]), rel=1e-10)
            np.testing.assert_allclose(fd.sample_llh(), gold[f"{tag}/sample_llh"][t], rtol=1e-10)
            fd.sync_host_arrays()
            mc, w2 = fd.hist()
            np.testing.assert_allclose(mc, gold[f"{tag}/mc"][t], rtol=1e-12, atol=1e-13)
            np.testing.assert_allclose(w2, gold[f"{tag}/w2"][t], rtol=1e-12, atol=1e-13)
        fd.close()


@pytest.mark.gpu
@pytest.mark.parametrize("build", ["float", "double"])
def test_adapter_binned_arm_over_the_reference_real_classes(build):
    """The binned arm of SetSplinePointers (Samples/SampleHandlerFD.cpp:1196-1242): the reference's REAL
    BinnedSplineHandler + SampleHandlerFD wired by the reference's harness, in BOTH builds of the reference
    (_LOW_MEMORY_STRUCTS_: M3::float_t = float; default: double -- the adapter's types follow M3::float_t); MoveToB200 turns
    every `&weightvec_Monolith[slot]` pointer into a slot index and the step runs on the B200 (binned_eval_kernel +
    binned_fill_kernel).  -lnL and histograms against the reference's own CPU results (tests/golden/ref_host_fd.npz)."""
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    import refpath_cases as RC
    from mach3_b200.synth import binned as B
    from oracle import ref_path_binding as RP
    if not RP.available_b200(f"{build}_b200"):
        pytest.skip(f"oracle/_ref adapter library of the {build} build not built (needs /root/reference at build time)")
    gold = np.load(os.path.join(here, "golden", "ref_host_fd.npz"))
    tag = f"binned_{build}"
    f64 = build == "double"
    w = RC.binned_workload()
    spl, ev = B.make_binned_splines(w, f64=f64), B.make_binned_events(w, f64=f64)
    E = w.n_events
    fd = RP.RefSampleHandlerFD(B.bin_edges(w), 1, True, build=f"{build}_b200")
    fd.attach_binned(spl)
    idx = np.arange(E, dtype=np.int32)
    fd.set_events(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, w.n_norm_params, w_before=idx,
                  w_after=E + idx, n_pool=2 * E, binned_n_per_event=ev["n_per_event"], binned_slot=ev["spline_index"])
    # one step on the reference's CPU path first (the adapter forwards to the base class before MoveToB200): the weight
    # pool -- the constant extra weights the adapter folds into the events' static weights -- holds its values from here on
    sp, nm = B.proposal(w, RC.BINNED_STEPS[0])
    fd.reweight(sp, nm, np.concatenate([B.make_osc(w, max(RC.BINNED_STEPS[0], 0), f64=f64), ev["static_w"]]).astype(np.float64))
    fd.set_data(gold[f"{tag}/data"])
    fd.move_to_b200(0)
    for i, step in enumerate(RC.BINNED_STEPS):
        sp, nm = B.proposal(w, step)
        pool = np.concatenate([B.make_osc(w, max(step, 0), f64=f64), ev["static_w"]]).astype(np.float64)
        fd.reweight(sp, nm, pool)
        assert fd.llh() == pytest.approx(float(gold[f"{tag}/llh"][i]), rel=1e-10), step
        fd.sync_host_arrays()
        mc, w2 = fd.hist()
        np.testing.assert_allclose(mc, gold[f"{tag}/mc"][i], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(w2, gold[f"{tag}/w2"][i], rtol=1e-12, atol=1e-13)
    fd.close()


@pytest.mark.gpu
def test_reference_smonolith_cuda_build_drives_the_drop_in_smonolithgpu():
    """Level 1 of INTEGRATION.md for real: the reference's SMonolith compiled with MaCh3_CUDA (from
    /root/reference/Splines/SplineMonolith.cpp) and linked with adapters/SMonolithGPU_m3b200.cu in place of the
    reference's gpuSplineUtils.cu.  Its constructor (MoveToGPU) and Evaluate() go through the reference's own
    SMonolithGPU interface into libm3b200; the per-event weights equal, bit for bit, what the same class computes in
    its CPU build (tests/golden/ref_host_path.npz), and SampleHandlerFD on top gives the reference's histograms."""
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    import refpath_cases as RC
    from oracle import ref_path_binding as RP
    if not RP.available_cuda():
        pytest.skip("oracle/_ref/libm3ref_path_lm_cuda.so not built (needs /root/reference at build time)")
    gold = np.load(os.path.join(here, "golden", "ref_host_path.npz"))
    for name in RC.CASES:
        c = RC.make_case(name)
        m = RP.RefSMonolith(c["type"], c["npts"], c["vals"], build="float_cuda")
        # float build: FastSplineInfo::xPts are floats, so segments can differ from the double-build vectors within a
        # float ulp of a knot (tests/test_reference_path.py); compare the steps where they agree
        for t in range(c["pars"].shape[0]):
            w, s, v = m.evaluate(c["pars"][t])
            np.testing.assert_array_equal(v, gold[f"{name}/param_values"][t])
            if np.array_equal(s, gold[f"{name}/segments"][t]):
                np.testing.assert_array_equal(w.view(np.uint32), gold[f"{name}/weights"][t].view(np.uint32), err_msg=f"{name} step {t}")
            else:
                assert t in (5, 9, 13, 17, 21, 25, 33)
        m.close()
    # and the full sample handler on top of the GPU-built monolith: the reference's FillArray reads the weights the
    # drop-in class copied back into cpu_total_weights
    gfd = np.load(os.path.join(here, "golden", "ref_host_fd.npz"))
    f = RC.fd_case()
    c = f["mono"]
    E = f["sample_id"].size
    m = RP.RefSMonolith(c["type"], c["npts"], c["vals"], build="float_cuda")
    fd = RP.RefSampleHandlerFD(RC.fd_edges(), 1, True, build="float_cuda")
    fd.attach_monolith(m)
    idx = np.arange(E, dtype=np.int32)
    fd.set_events(f["sample_id"], f["kin"], f["norm_idx"], RC.NPE, RC.N_NORM, w_before=idx, w_after=E + idx, n_pool=2 * E)
    for t in range(8):
        pool = np.concatenate([f["osc"][t], f["static_w"]]).astype(np.float64)
        fd.reweight(f["pars"][t], f["norm"][t], pool)
        if t == 0:
            fd.set_data(gfd["mono_w21/data"])
        mc, w2 = fd.hist()
        np.testing.assert_array_equal(mc, gfd["mono_w21/mc"][t])          # same weights, same serial FillArray
        assert fd.llh() == pytest.approx(float(gfd["mono_w21/llh"][t]), rel=1e-14)
    fd.close()
