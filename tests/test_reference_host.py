"""Pins the binning part of the path to the REFERENCE'S OWN CODE: struct SampleBinningInfo / BinInfo
(Samples/SampleStructs.h: FindBin :577-613, InitialiseBinMigrationLookUp :618-675, InitNonUniform +
InitialiseGridMapping :394-528, IsEventInside :207-219) was compiled from /root/reference with empty stand-ins for
the absent ROOT/spdlog headers (oracle/ref_host) and run on edge-heavy inputs; its outputs are committed as
tests/golden/ref_host_binning.npz (generator: tests/golden/make_ref_host_golden.py).
  * CPU: the oracle's restatement must reproduce them bit-for-bit (bin ids, mega-grid edges, box lists);
  * live: where oracle/_ref/libm3ref_host.so exists, the reference is re-run and must reproduce the vectors;
  * GPU: the device's bin ids (bin_kernel, through the C ABI) must equal the reference's."""
import os

import numpy as np
import pytest

from oracle import binding as O
from oracle import ref_host_binding as RH

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_host_binning.npz")
CASES = ["uni1d", "uni2d", "uni3d", "box_a", "box_b", "box_doc"]


def _spec(g, name):
    s = g[f"{name}/spec"]
    if s.ndim == 3:
        return s
    d, out = 0, []
    while f"{name}/in_edges{d}" in g:
        out.append(g[f"{name}/in_edges{d}"]); d += 1
    return out


@pytest.mark.parametrize("name", CASES)
def test_oracle_binning_matches_reference_vectors(name):
    g = np.load(GOLD)
    spec = _spec(g, name)
    kin = g[f"{name}/kin"]
    nd, n = kin.shape
    sh = O.SampleHandlerFD(n, [spec])
    assert sh.n_bins == int(g[f"{name}/n_bins"][0])
    for d in range(nd):
        np.testing.assert_array_equal(sh.axis_edges(0, d), g[f"{name}/edges{d}"])          # same doubles, incl. the mega grid's
        for key in ("true", "pert"):
            nom = g[f"{name}/nom_{key}"][d]
            got = np.array([sh.find_bin(0, d, kin[d, i], int(nom[i])) for i in range(n)], np.int32)
            np.testing.assert_array_equal(got, g[f"{name}/findbin_{key}{d}"])
    sh.set_events(np.zeros(n, np.int32), kin.reshape(-1), None, 0, None, None, None, None)
    np.testing.assert_array_equal(sh.event_bins(), g[f"{name}/bin_true"])
    if isinstance(spec, np.ndarray):
        lens, idx = g[f"{name}/grid_len"], g[f"{name}/grid_idx"]
        off = 0
        for mega, ln in enumerate(lens):
            assert sh.grid_mapping(0, mega) == idx[off:off + ln].tolist()
            off += ln
    # perturbed nominal bins never change the answer (the shortcuts agree with the binary search)
    np.testing.assert_array_equal(g[f"{name}/bin_true"], g[f"{name}/bin_pert"])


@pytest.mark.skipif(not RH.available(), reason="oracle/_ref/libm3ref_host.so not built (needs /root/reference at build time)")
@pytest.mark.parametrize("name", CASES)
def test_reference_rerun_reproduces_the_vectors(name):
    g = np.load(GOLD)
    ref = RH.RefBinning(_spec(g, name))
    kin = g[f"{name}/kin"]
    np.testing.assert_array_equal(ref.find_sample_bin(kin, g[f"{name}/nom_true"]), g[f"{name}/bin_true"])
    np.testing.assert_array_equal(ref.find_sample_bin(kin, g[f"{name}/nom_pert"]), g[f"{name}/bin_pert"])
    ref.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_bins_match_reference_vectors(name):
    from mach3_b200 import lib
    g = np.load(GOLD)
    spec = _spec(g, name)
    kin = g[f"{name}/kin"]
    h = lib.Handle()
    h.upload_binning([spec])
    h.upload_events(np.zeros(kin.shape[1], np.int32), np.ascontiguousarray(kin).reshape(-1))
    np.testing.assert_array_equal(h.read_event_bins(), g[f"{name}/bin_true"])
    h.close()


from oracle import ref_path_binding as RP   # noqa: E402


@pytest.mark.skipif(not RP.available(), reason="oracle/_ref/libm3ref_path.so not built (needs /root/reference at build time)")
@pytest.mark.parametrize("name", CASES)
def test_real_binninghandler_findglobalbin_reproduces_the_vectors(name):
    """The vectors above were made with the ten glue lines of BinningHandler::FindGlobalBin re-stated in the harness
    (Samples/BinningHandler.cpp could not be compiled then).  It compiles now (oracle/ref_host/harness_path.cpp):
    the reference's REAL BinningHandler::FindGlobalBin, with the nominal bins from its own FindNominalBinAndEdges,
    gives the same bin ids on the same edge-heavy inputs."""
    g = np.load(GOLD)
    spec = _spec(g, name)
    kin = g[f"{name}/kin"]
    fd = RP.RefSampleHandlerFD([spec], 0, False, build="double")
    fd.set_events(np.zeros(kin.shape[1], np.int32), kin)
    _, bins = fd.events()
    np.testing.assert_array_equal(bins, g[f"{name}/bin_true"])
    fd.close()
