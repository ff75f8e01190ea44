"""Generates tests/golden/ref_host_selection.npz: the REFERENCE's own SampleHandlerFD::IsEventSelected
(Samples/SampleHandlerFD.cpp:281-294) inside its FillArray (:352-386), compiled from /root/reference by
oracle/ref_host/Makefile (libm3ref_path_lm.so, the float build that wires SMonolith into SampleHandlerFD) and run
HERE on the seeded inputs of tests/refpath_cases.py with a non-empty StoredSelection (refpath_cases.selection_case).

    python tests/golden/make_ref_host_selection.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refpath_cases as RC                      # noqa: E402
from oracle import ref_path_binding as RP      # noqa: E402

BARLOW_BEESTON = 1


def main():
    f = RC.fd_case()
    sel = RC.selection_case()
    c = f["mono"]
    E = f["sample_id"].size
    out = {}
    for update_w2 in (False, True):
        tag = f"sel_w2{int(update_w2)}"
        m = RP.RefSMonolith(c["type"], c["npts"], c["vals"], build="float")
        fd = RP.RefSampleHandlerFD(RC.fd_edges(), BARLOW_BEESTON, update_w2, build="float")
        fd.attach_monolith(m)
        idx = np.arange(E, dtype=np.int32)
        fd.set_events(f["sample_id"], sel["kin4"], f["norm_idx"], RC.NPE, RC.N_NORM, w_before=idx, w_after=E + idx, n_pool=2 * E)
        fd.set_selection(sel["cuts"])
        rec = {k: [] for k in ("mc", "w2", "llh", "sample_llh", "selected", "event_bin")}
        for t in range(RC.SEL_STEPS):
            if t == RC.SEL_SHIFT_AT:
                fd.set_kin(sel["kin4_shift"])                # functional shifts move binning AND cut variables
            pool = np.concatenate([f["osc"][t], f["static_w"]]).astype(np.float64)
            fd.reweight(f["pars"][t], f["norm"][t], pool)
            if t == 0:
                mc, _ = fd.hist()
                data = np.random.default_rng(33).poisson(mc).astype(np.float64)
                fd.set_data(data)
                out[f"{tag}/data"] = data
            mc, w2 = fd.hist()
            _, b = fd.events()
            rec["mc"].append(mc); rec["w2"].append(w2); rec["llh"].append(fd.llh()); rec["sample_llh"].append(fd.sample_llh())
            rec["selected"].append(fd.selected()); rec["event_bin"].append(b)
        for k, v in rec.items():
            out[f"{tag}/{k}"] = np.asarray(v)
        print(tag, "selected", [int(s.sum()) for s in rec["selected"]], "of", E, "llh", rec["llh"][:3])
        fd.close()
    path = os.path.join(ROOT, "tests", "golden", "ref_host_selection.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
