"""SampleHandlerFD::IsEventSelected (Samples/SampleHandlerFD.cpp:281-294; applied at :361 / :424, after ApplyShifts and
before CalcWeightTotal) -- kinematic selection cuts.

Golden vectors: tests/golden/ref_host_selection.npz, produced by the REFERENCE'S OWN SampleHandlerFD compiled from
/root/reference (generator tests/golden/make_ref_host_selection.py) with a non-empty StoredSelection on the three
samples of refpath_cases.fd_edges(): cuts on binning variables, on cut-only variables, values exactly on both bounds
(lower passes, upper fails), NaN (passes), and shifted kinematics from step SEL_SHIFT_AT on.
  * CPU: the oracle reproduces the reference's selected mask, histograms and -lnL bit for bit;
  * live: where oracle/_ref/libm3ref_path_lm.so exists the reference is re-run and must reproduce the vectors;
  * GPU: libm3b200 (m3b_upload_selection / m3b_update_selection_values) gives the same mask bit for bit, histograms
    to 1e-12, -lnL to 1e-10; removing the selection restores the unselected result."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refpath_cases as RC  # noqa: E402
from oracle import binding as O
from oracle import ref_path_binding as RP

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "ref_host_path.npz")
GOLD_SEL = os.path.join(HERE, "golden", "ref_host_selection.npz")
GOLD_FD = os.path.join(HERE, "golden", "ref_host_fd.npz")
BARLOW_BEESTON = 1
ARR = ("coeff_x", "coeff_many", "nKnots_arr", "paramNo_arr", "nParamPerEvent", "nParamPerEvent_tf1", "paramNo_tf1",
       "coeff_tf1", "n_pts")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def gold_sel():
    return np.load(GOLD_SEL)


def _arrays(g):
    a = {k: g[f"mixed/arr/{k}"] for k in ARR}
    a["n_events"] = int(g["mixed/sizes"][0])
    return a, int(g["mixed/sizes"][1]), int(g["mixed/sizes"][2])


@pytest.fixture
def serial_oracle():
    was = O.lib().m3o_get_multithread()
    O.set_multithread(False)
    yield
    O.set_multithread(bool(was))


@pytest.mark.parametrize("update_w2", [False, True])
def test_oracle_selection_matches_reference(gold, gold_sel, update_w2, serial_oracle):
    tag = f"sel_w2{int(update_w2)}"
    f, sel = RC.fd_case(), RC.selection_case()
    E = f["sample_id"].size
    a, P, K = _arrays(gold)
    mono = O.SMonolith(P, K, a["coeff_x"], a["n_pts"], a)
    sh = O.SampleHandlerFD(E, RC.fd_edges(), BARLOW_BEESTON, update_w2)
    norm, osc = np.ones(RC.N_NORM), np.ones(E, np.float32)
    sh.set_events(f["sample_id"], sel["kin4"][:2].reshape(-1).copy(), f["norm_idx"].reshape(-1), RC.NPE, norm, osc, mono, f["static_w"])
    sh.SetSelection(sel["cuts"], sel["kin4"].copy())
    for t in range(RC.SEL_STEPS):
        if t == RC.SEL_SHIFT_AT:
            sh._keep[1][:] = sel["kin4_shift"][:2].reshape(-1)       # what the KinVar pointers look at
            sh.cut_values[:] = sel["kin4_shift"]                     # what ReturnKinematicParameter returns
        mono.set_params(f["pars"][t]); sh.norm_vals[:] = f["norm"][t]; sh.osc_w[:] = f["osc"][t]
        sh.Reweight()
        if t == 0:
            sh.AddData(gold_sel[f"{tag}/data"])
        np.testing.assert_array_equal(sh.event_selected(), gold_sel[f"{tag}/selected"][t], err_msg=f"selected, step {t}")
        np.testing.assert_array_equal(sh.event_bins(), gold_sel[f"{tag}/event_bin"][t])
        np.testing.assert_array_equal(sh.mc, gold_sel[f"{tag}/mc"][t], err_msg=f"mc, step {t}")
        np.testing.assert_array_equal(sh.w2, gold_sel[f"{tag}/w2"][t], err_msg=f"w2, step {t}")
        assert sh.GetLikelihood() == pytest.approx(float(gold_sel[f"{tag}/llh"][t]), rel=1e-14)


def test_selection_vectors_exercise_the_edge_cases(gold_sel):
    """The committed vectors really contain what the docstring promises."""
    f, sel = RC.fd_case(), RC.selection_case()
    s1 = f["sample_id"] == 1
    v3 = sel["kin4"][3]
    got = gold_sel["sel_w20/selected"][0]
    assert (v3[s1] == 0.25).any() and got[s1 & (v3 == 0.25)].all()            # Val == LowerBound passes
    assert (v3[s1] == 0.75).any() and not got[s1 & (v3 == 0.75)].any()        # Val == UpperBound fails
    assert np.isnan(v3[s1]).any() and got[s1 & np.isnan(v3)].all()            # NaN: neither comparison is true
    assert 0 < got.sum() < got.size
    assert (gold_sel["sel_w20/selected"][0] != gold_sel["sel_w20/selected"][RC.SEL_SHIFT_AT]).any()
    # the cuts matter: the unselected histograms of the same inputs differ
    assert not np.array_equal(gold_sel["sel_w20/mc"][1], np.load(GOLD_FD)["mono_w20/mc"][1])


@pytest.mark.skipif(not RP.available(), reason="oracle/_ref/libm3ref_path_lm.so not built (needs /root/reference at build time)")
def test_reference_rerun_reproduces_the_selection_vectors(gold_sel, tmp_path, monkeypatch):
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_ref_host_selection as G
    monkeypatch.setattr(G, "ROOT", str(tmp_path))
    os.makedirs(tmp_path / "tests" / "golden")
    G.main()
    again = np.load(tmp_path / "tests" / "golden" / "ref_host_selection.npz")
    assert set(again.files) == set(gold_sel.files)
    for k in again.files:
        np.testing.assert_array_equal(again[k], gold_sel[k], err_msg=k)


def _device(gold, update_w2, f, sel, **kw):
    from mach3_b200 import handlers
    E = f["sample_id"].size
    a, P, K = _arrays(gold)
    sh = handlers.SampleHandlerFD(RC.fd_edges(), BARLOW_BEESTON, update_w2, keep_event_weights=True, keep_kinematics=True, **kw)
    sh.SetupSplines(P, K, a["coeff_x"], a["n_pts"], a)
    pars, norm, osc = np.zeros(P), np.ones(RC.N_NORM), np.ones(E, np.float32)
    sh.SetupEvents(f["sample_id"], sel["kin4"][:2].reshape(-1), f["norm_idx"].reshape(-1), RC.NPE, norm, osc, None, f["static_w"])
    sh.SetSplinePointers(pars)
    return sh, pars, norm, osc


@pytest.mark.gpu
@pytest.mark.parametrize("update_w2", [False, True])
@pytest.mark.parametrize("by_kin", [False, True])
def test_device_selection_matches_reference(gold, gold_sel, update_w2, by_kin):
    """by_kin: the cuts on binning variables refer to the kinematics the library already holds (cut_var = -1-d), so
    m3b_update_kinematics alone moves them; otherwise every cut variable comes through the caller's table."""
    tag = f"sel_w2{int(update_w2)}"
    f, sel = RC.fd_case(), RC.selection_case()
    sh, pars, norm, osc = _device(gold, update_w2, f, sel)
    cuts = [(s, (-1 - v if (by_kin and v < 2) else v), lo, hi) for s, v, lo, hi in sel["cuts"]]
    sh.SetSelection(cuts, sel["kin4"])
    for t in range(RC.SEL_STEPS):
        if t == RC.SEL_SHIFT_AT:
            sh.handle.update_kinematics(sel["kin4_shift"][:2].reshape(-1))
            sh.handle.update_selection_values(sel["kin4_shift"])
        pars[:] = f["pars"][t]; norm[:] = f["norm"][t]; osc[:] = f["osc"][t]
        sh.OscillatorEvaluated()
        sh.Reweight()
        if t == 0:
            sh.GetLikelihood()
            sh.AddData(gold_sel[f"{tag}/data"])
        llh = sh.GetLikelihood()
        np.testing.assert_array_equal(sh.handle.read_event_selected(), gold_sel[f"{tag}/selected"][t], err_msg=f"selected, step {t}")
        np.testing.assert_array_equal(sh.handle.read_event_bins(), gold_sel[f"{tag}/event_bin"][t], err_msg=f"bins, step {t}")
        mc, w2 = sh.handle.read_hist()
        np.testing.assert_allclose(mc, gold_sel[f"{tag}/mc"][t], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(w2, gold_sel[f"{tag}/w2"][t], rtol=1e-12, atol=1e-13)
        if t > 0:
            assert llh == pytest.approx(float(gold_sel[f"{tag}/llh"][t]), rel=1e-10)
            per = sh.handle.llh(per_sample=True)[1]
            np.testing.assert_allclose(per, gold_sel[f"{tag}/sample_llh"][t], rtol=1e-10)


@pytest.mark.gpu
def test_device_selection_can_be_removed_and_applies_to_batches(gold, gold_sel):
    """n_cuts = 0 restores the unselected fill (the reference's golden vectors without cuts); the batched kernel
    (m3b_step_batch) sees the same selection as the single-set kernel."""
    gfd = np.load(GOLD_FD)
    f, sel = RC.fd_case(), RC.selection_case()
    sh, pars, norm, osc = _device(gold, False, f, sel)
    sh.SetSelection(sel["cuts"], sel["kin4"])
    pars[:] = f["pars"][0]; norm[:] = f["norm"][0]; osc[:] = f["osc"][0]
    sh.OscillatorEvaluated(); sh.Reweight(); sh.GetLikelihood()
    sh.AddData(gold_sel["sel_w20/data"])
    # batch of the next three proposals (W2 is frozen now): each -lnL = the reference's with the cuts
    sh.handle.upload_osc(f["osc"][1])
    tot = sh.handle.step_batch(f["pars"][1:2], f["norm"][1:2])
    assert tot[0] == pytest.approx(float(gold_sel["sel_w20/llh"][1]), rel=1e-10)
    sh.SetSelection([], None)
    assert sh.handle.read_event_selected().all()
    pars[:] = f["pars"][1]; norm[:] = f["norm"][1]; osc[:] = f["osc"][1]
    sh.OscillatorEvaluated(); sh.Reweight(); sh.GetLikelihood()
    np.testing.assert_allclose(sh.handle.read_hist()[0], gfd["mono_w20/mc"][1], rtol=1e-12, atol=1e-13)


@pytest.mark.gpu
def test_selection_argument_errors(gold):
    from mach3_b200 import handlers, lib
    f, sel = RC.fd_case(), RC.selection_case()
    a, P, K = _arrays(gold)
    sh = handlers.SampleHandlerFD(RC.fd_edges(), 0, False)
    with pytest.raises(lib.M3BError):                                  # before the events
        sh.handle.n_events = f["sample_id"].size
        sh.SetSelection(sel["cuts"], sel["kin4"])
    E = f["sample_id"].size
    sh.SetupEvents(f["sample_id"], sel["kin4"][:2].reshape(-1), None, 0, None, np.ones(E, np.float32), None, None)
    with pytest.raises(lib.M3BError):                                  # sample out of range
        sh.SetSelection([(7, 0, 0.0, 1.0)], sel["kin4"])
    with pytest.raises(lib.M3BError):                                  # variable out of range
        sh.SetSelection([(0, 4, 0.0, 1.0)], sel["kin4"])
    with pytest.raises(lib.M3BError):                                  # binning variable without KEEP_KINEMATICS
        sh.SetSelection([(0, -1, 0.0, 1.0)], sel["kin4"])
    sh.SetSelection(sel["cuts"], sel["kin4"])                          # and the handle still works
