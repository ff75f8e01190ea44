"""Functional ("shift") parameters: SampleHandlerFD::ApplyShifts (Samples/SampleHandlerFD.cpp:545-564), applied per event
and per step before IsEventSelected and FindGlobalBin.

Golden vectors: tests/golden/ref_host_shifts.npz -- the REFERENCE'S OWN ApplyShifts loop (ResetShifts -> every
FunctionalShifter of funcParsGrid[event] -> FinaliseShifts) compiled from /root/reference and driven with three linear
functional parameters (refpath_cases.shift_case: a scale of a binning variable, an offset on the other, a move of a
cut-only variable), generator tests/golden/make_ref_host_shifts.py.
  * CPU: the oracle reproduces shifted kinematics, selection mask, bins, histograms and -lnL bit for bit;
  * live: the reference is re-run where oracle/_ref/libm3ref_path_lm.so exists;
  * GPU: m3b_upload_linear_shifts + m3b_set_shift_pars (shift_kernel -> bin_kernel -> select_kernel, nothing per-event
    crosses PCIe) gives the same mask and bins bit for bit, histograms to 1e-12, -lnL to 1e-10 -- single handle and a
    three-member group."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refpath_cases as RC  # noqa: E402
from oracle import binding as O
from oracle import ref_path_binding as RP

HERE = os.path.dirname(os.path.abspath(__file__))
BARLOW_BEESTON = 1
ARR = ("coeff_x", "coeff_many", "nKnots_arr", "paramNo_arr", "nParamPerEvent", "nParamPerEvent_tf1", "paramNo_tf1",
       "coeff_tf1", "n_pts")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "ref_host_path.npz"))


@pytest.fixture(scope="module")
def gold_sh():
    return np.load(os.path.join(HERE, "golden", "ref_host_shifts.npz"))


def _arrays(g):
    a = {k: g[f"mixed/arr/{k}"] for k in ARR}
    a["n_events"] = int(g["mixed/sizes"][0])
    return a, int(g["mixed/sizes"][1]), int(g["mixed/sizes"][2])


def test_oracle_shifts_match_reference(gold, gold_sh):
    was = O.lib().m3o_get_multithread()
    O.set_multithread(False)
    f, sel, sh = RC.fd_case(), RC.selection_case(), RC.shift_case()
    E = f["sample_id"].size
    a, P, K = _arrays(gold)
    mono = O.SMonolith(P, K, a["coeff_x"], a["n_pts"], a)
    osh = O.SampleHandlerFD(E, RC.fd_edges(), BARLOW_BEESTON, False)
    norm, osc = np.ones(RC.N_NORM), np.ones(E, np.float32)
    osh.set_events(f["sample_id"], sel["kin4"][:2].reshape(-1).copy(), f["norm_idx"].reshape(-1), RC.NPE, norm, osc, mono, f["static_w"])
    osh.SetSelection(sel["cuts"], sel["kin4"].copy())
    npe, par, tgt, cf = RC.shift_entries(sh)
    osh.SetLinearShifts(npe, par, tgt, cf, np.zeros(RC.N_SHIFT))
    for t in range(RC.SHIFT_STEPS):
        osh.shift_values[:] = sh["values"][t]
        mono.set_params(f["pars"][t]); osh.norm_vals[:] = f["norm"][t]; osh.osc_w[:] = f["osc"][t]
        osh.Reweight()
        if t == 0:
            osh.AddData(gold_sh["shift/data"])
        # shifted kinematics as the reference left them (events without shifters keep whatever they had: nominal)
        kin = osh._keep[1].reshape(2, E)
        np.testing.assert_array_equal(kin[0], gold_sh["shift/kin"][t][:, 0], err_msg=f"kin 0, step {t}")
        np.testing.assert_array_equal(kin[1], gold_sh["shift/kin"][t][:, 1], err_msg=f"kin 1, step {t}")
        np.testing.assert_array_equal(osh.cut_values[2], gold_sh["shift/kin"][t][:, 2], err_msg=f"cut variable 2, step {t}")
        np.testing.assert_array_equal(osh.event_selected(), gold_sh["shift/selected"][t], err_msg=f"selected, step {t}")
        np.testing.assert_array_equal(osh.event_bins(), gold_sh["shift/event_bin"][t], err_msg=f"bins, step {t}")
        np.testing.assert_array_equal(osh.mc, gold_sh["shift/mc"][t], err_msg=f"mc, step {t}")
        assert osh.GetLikelihood() == pytest.approx(float(gold_sh["shift/llh"][t]), rel=1e-14)
    O.set_multithread(bool(was))


def test_shift_vectors_move_events(gold_sh):
    assert (gold_sh["shift/event_bin"][0] != gold_sh["shift/event_bin"][1]).any()
    assert (gold_sh["shift/selected"][0] != gold_sh["shift/selected"][3]).any()
    assert not np.array_equal(gold_sh["shift/kin"][0], gold_sh["shift/kin"][2])


@pytest.mark.skipif(not RP.available(), reason="oracle/_ref/libm3ref_path_lm.so not built (needs /root/reference at build time)")
def test_reference_rerun_reproduces_the_shift_vectors(gold_sh, tmp_path, monkeypatch):
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_ref_host_shifts as G
    monkeypatch.setattr(G, "ROOT", str(tmp_path))
    os.makedirs(tmp_path / "tests" / "golden")
    G.main()
    again = np.load(tmp_path / "tests" / "golden" / "ref_host_shifts.npz")
    assert set(again.files) == set(gold_sh.files)
    for k in again.files:
        np.testing.assert_array_equal(again[k], gold_sh[k], err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize("n_members", [1, 3])
def test_device_shifts_match_reference(gold, gold_sh, n_members):
    import torch
    from mach3_b200 import lib
    f, sel, sh = RC.fd_case(), RC.selection_case(), RC.shift_case()
    E = f["sample_id"].size
    a, P, K = _arrays(gold)
    spl = dict(a); spl["nKnots_arr"] = np.asarray(a["nKnots_arr"], np.uint32)
    npe, par, tgt, cf = RC.shift_entries(sh)
    flags = lib.FLAG_KEEP_KINEMATICS
    if n_members == 1:
        x = lib.Handle(test_statistic=BARLOW_BEESTON, update_w2=False, flags=flags)
    else:
        x = lib.Group([i % torch.cuda.device_count() for i in range(n_members)], test_statistic=BARLOW_BEESTON, update_w2=False, flags=flags)
    x.upload_binning(RC.fd_edges())
    x.upload_spline_monolith(P, K, a["coeff_x"], a["n_pts"], spl)
    x.upload_events(f["sample_id"], sel["kin4"][:2].reshape(-1), f["norm_idx"].reshape(-1), RC.NPE, RC.N_NORM, True, None, 0, f["static_w"])
    x.upload_selection(sel["cuts"], sel["kin4"])
    x.upload_linear_shifts(RC.N_SHIFT, npe, par, tgt, cf)
    if n_members > 1:
        x.connect("peer")
    for t in range(RC.SHIFT_STEPS):
        x.set_shift_pars(sh["values"][t])
        x.upload_osc(f["osc"][t])
        x.step(f["pars"][t], f["norm"][t])
        llh = x.llh()
        if t == 0:
            x.upload_data(gold_sh["shift/data"])
            x.step(f["pars"][t], f["norm"][t])
            llh = x.llh()
        if n_members == 1:
            np.testing.assert_array_equal(x.read_event_selected(), gold_sh["shift/selected"][t], err_msg=f"selected, step {t}")
            np.testing.assert_array_equal(x.read_event_bins(), gold_sh["shift/event_bin"][t], err_msg=f"bins, step {t}")
        mc, _ = x.read_hist()
        np.testing.assert_allclose(mc, gold_sh["shift/mc"][t], rtol=1e-12, atol=1e-13)
        assert llh == pytest.approx(float(gold_sh["shift/llh"][t]), rel=1e-10)
    x.close()
