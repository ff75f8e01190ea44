"""Generates tests/golden/ref_host_fd.npz: outputs of the REFERENCE's own SampleHandlerFD::Reweight / GetLikelihood
(Samples/SampleHandlerFD.cpp), BinningHandler::FindGlobalBin (Samples/BinningHandler.cpp), SMonolith::Evaluate and
BinnedSplineHandler::Evaluate (Splines/*.cpp), compiled from /root/reference by oracle/ref_host/Makefile in both of
the reference's builds (libm3ref_path.so: M3::float_t = double; libm3ref_path_lm.so: _LOW_MEMORY_STRUCTS_, float)
and run HERE on the seeded inputs of tests/refpath_cases.py.

    python tests/golden/make_ref_host_fd.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refpath_cases as RC                      # noqa: E402
from oracle import ref_path_binding as RP      # noqa: E402
from mach3_b200.synth import binned as B       # noqa: E402

BARLOW_BEESTON = 1


def run_monolith(out, update_w2):
    tag = f"mono_w2{int(update_w2)}"
    f = RC.fd_case()
    c = f["mono"]
    E = f["sample_id"].size
    m = RP.RefSMonolith(c["type"], c["npts"], c["vals"], build="float")
    fd = RP.RefSampleHandlerFD(RC.fd_edges(), BARLOW_BEESTON, update_w2, build="float")
    fd.attach_monolith(m)
    idx = np.arange(E, dtype=np.int32)
    fd.set_events(f["sample_id"], f["kin"], f["norm_idx"], RC.NPE, RC.N_NORM, w_before=idx, w_after=E + idx, n_pool=2 * E)
    rec = {k: [] for k in ("mc", "w2", "llh", "sample_llh", "event_w", "event_bin", "segments")}
    for t in range(RC.FD_STEPS):
        if t == 8:
            fd.set_kin(f["kin_shift"])                       # functional shifts: new kinematics, same nominal bins
        pool = np.concatenate([f["osc"][t], f["static_w"]]).astype(np.float64)
        fd.reweight(f["pars"][t], f["norm"][t], pool)
        if t == 0:
            mc, _ = fd.hist()
            data = np.random.default_rng(31).poisson(mc).astype(np.float64)
            fd.set_data(data)
            out[f"{tag}/data"] = data
        mc, w2 = fd.hist()
        w, b = fd.events()
        rec["mc"].append(mc); rec["w2"].append(w2); rec["llh"].append(fd.llh()); rec["sample_llh"].append(fd.sample_llh())
        rec["event_w"].append(w.astype(np.float32)); rec["event_bin"].append(b); rec["segments"].append(fd.segments(c["pars"].shape[1]))
    for k, v in rec.items():
        out[f"{tag}/{k}"] = np.asarray(v)
    out[f"{tag}/n_bins"] = np.array([fd.n_bins])
    print(tag, "bins", fd.n_bins, "llh", rec["llh"][:3])
    fd.close()


def run_binned(out, build):
    tag = f"binned_{build}"
    w = RC.binned_workload()
    f64 = build == "double"
    spl = B.make_binned_splines(w, f64=f64)
    ev = B.make_binned_events(w, f64=f64)
    E = w.n_events
    fd = RP.RefSampleHandlerFD(B.bin_edges(w), BARLOW_BEESTON, True, build=build)
    fd.attach_binned(spl)
    idx = np.arange(E, dtype=np.int32)
    fd.set_events(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, w.n_norm_params, w_before=idx,
                  w_after=E + idx, n_pool=2 * E, binned_n_per_event=ev["n_per_event"], binned_slot=ev["spline_index"])
    rec = {k: [] for k in ("mc", "w2", "llh", "slot_w", "event_w", "segments")}
    for i, step in enumerate(RC.BINNED_STEPS):
        sp, nm = B.proposal(w, step)
        pool = np.concatenate([B.make_osc(w, max(step, 0), f64=f64), ev["static_w"]]).astype(np.float64)
        fd.reweight(sp, nm, pool)
        if i == 0:
            mc, _ = fd.hist()
            data = np.random.default_rng(32).poisson(mc).astype(np.float64)
            fd.set_data(data)
            out[f"{tag}/data"] = data
        mc, w2 = fd.hist()
        ew, _ = fd.events()
        rec["mc"].append(mc); rec["w2"].append(w2); rec["llh"].append(fd.llh()); rec["slot_w"].append(fd.binned_weights())
        rec["event_w"].append(ew); rec["segments"].append(fd.segments(w.n_systs))
    for k, v in rec.items():
        a = np.asarray(v)
        out[f"{tag}/{k}"] = a.astype(np.float32) if (build == "float" and k in ("slot_w", "event_w")) else a
    print(tag, "llh", rec["llh"][:3])
    fd.close()


def main():
    out = {}
    run_monolith(out, False)
    run_monolith(out, True)
    run_binned(out, "float")
    run_binned(out, "double")
    path = os.path.join(ROOT, "tests", "golden", "ref_host_fd.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
