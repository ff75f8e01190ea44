"""Generates tests/golden/ref_host_shifts.npz: the REFERENCE's own SampleHandlerFD::ApplyShifts
(Samples/SampleHandlerFD.cpp:545-564: ResetShifts -> every FunctionalShifter of funcParsGrid[event] -> FinaliseShifts)
inside its FillArray, with IsEventSelected and FindGlobalBin acting on the shifted variables -- compiled from
/root/reference by oracle/ref_host/Makefile (libm3ref_path_lm.so) and run HERE on refpath_cases.shift_case().

    python tests/golden/make_ref_host_shifts.py
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
    f, sel, sh = RC.fd_case(), RC.selection_case(), RC.shift_case()
    c = f["mono"]
    E = f["sample_id"].size
    out = {}
    m = RP.RefSMonolith(c["type"], c["npts"], c["vals"], build="float")
    fd = RP.RefSampleHandlerFD(RC.fd_edges(), BARLOW_BEESTON, False, build="float")
    fd.attach_monolith(m)
    idx = np.arange(E, dtype=np.int32)
    fd.set_events(f["sample_id"], sel["kin4"], f["norm_idx"], RC.NPE, RC.N_NORM, w_before=idx, w_after=E + idx, n_pool=2 * E)
    fd.set_selection(sel["cuts"])
    fd.set_linear_shifts(sh["target"], sh["coef"])
    rec = {k: [] for k in ("mc", "w2", "llh", "selected", "event_bin", "kin")}
    for t in range(RC.SHIFT_STEPS):
        fd.set_shift_pars(sh["values"][t])
        pool = np.concatenate([f["osc"][t], f["static_w"]]).astype(np.float64)
        fd.reweight(f["pars"][t], f["norm"][t], pool)
        if t == 0:
            mc, _ = fd.hist()
            data = np.random.default_rng(34).poisson(mc).astype(np.float64)
            fd.set_data(data)
            out["shift/data"] = data
        mc, w2 = fd.hist()
        _, b = fd.events()
        rec["mc"].append(mc); rec["w2"].append(w2); rec["llh"].append(fd.llh())
        rec["selected"].append(fd.selected()); rec["event_bin"].append(b); rec["kin"].append(fd.kin())
    for k, v in rec.items():
        out[f"shift/{k}"] = np.asarray(v)
    print("selected", [int(s.sum()) for s in rec["selected"]], "of", E, "llh", rec["llh"])
    path = os.path.join(ROOT, "tests", "golden", "ref_host_shifts.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    fd.close()


if __name__ == "__main__":
    main()
