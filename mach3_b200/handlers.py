"""Host-side mirror of the reference's call surface for the hot path, on top of the C ABI.

Names, argument meaning and call order follow the reference so the parity tests read like
reference code:

    SMonolith            Splines/SplineMonolith.h:13-137 (Evaluate, SynchroniseMemTransfer,
                         retPointer, setSplinePointers, GetNParams, GetName)
    SampleHandlerFD      Samples/SampleHandlerFD.h:21-413 + SampleHandlerBase.h:37-90 (Reweight,
                         GetLikelihood, GetSampleLikelihood, AddData, GetMCArray, GetW2Array,
                         GetDataArray, SetTestStatistic)

"Pointers" are numpy arrays the caller keeps alive and mutates in place, exactly like
pyMaCh3's ``set_param_value_array`` (python/splines.cpp:195-232).
"""
from __future__ import annotations

import numpy as np

from . import lib as _lib


class SMonolith:
    """Event-by-event spline monolith living on the B200.

    Stand-alone use (pyMaCh3 ``EventSplineMonolith`` style): ``Evaluate()`` then
    ``retPointer(event)`` / ``cpu_total_weights``.  Attached to a :class:`SampleHandlerFD` the
    evaluation is fused into ``Reweight()`` and per-event weights are only materialised on demand.
    """

    def __init__(self, n_params, max_knots, coeff_x, n_pts, spl, device=0, handle=None, chunk_events=None):
        self.n_params = int(n_params)
        self.n_events = int(spl["n_events"]) if "n_events" in spl else int(np.asarray(spl["nParamPerEvent"]).size // 2)
        self._standalone = handle is None
        if handle is None:
            handle = _lib.Handle(device=device, flags=_lib.FLAG_KEEP_EVENT_WEIGHTS)
        self.handle = handle
        if chunk_events is None:
            handle.splines_begin(n_params, max_knots, coeff_x, n_pts, self.n_events)
            handle.splines_append(spl)
            handle.splines_end()
        else:  # exercised by the tests: arbitrary chunking must give the same device tables
            handle.splines_begin(n_params, max_knots, coeff_x, n_pts, self.n_events)
            for e0 in range(0, self.n_events, chunk_events):
                handle.splines_append(slice_monolith(spl, e0, min(self.n_events, e0 + chunk_events)))
            handle.splines_end()
        if self._standalone:
            # one sample, one bin that holds every event: the fill is a by-product, the weights are the output
            handle.upload_binning([[np.array([-1.0, 1.0])]])
            handle.upload_events(np.zeros(self.n_events, np.int32), np.zeros(self.n_events, np.float64))
        self._pars = None
        self.cpu_total_weights = None

    # -- reference API -------------------------------------------------------------------------
    def GetName(self):
        return "SplineMonolith"

    def GetNParams(self):
        return self.n_params

    def setSplinePointers(self, spline_pars_array):
        """The doubles behind FastSplineInfo::splineParsPointer; read again at every Evaluate."""
        assert spline_pars_array.dtype == np.float64 and spline_pars_array.size == self.n_params
        self._pars = spline_pars_array

    def Evaluate(self):
        if not self._standalone:
            raise RuntimeError("attached SMonolith is evaluated by SampleHandlerFD.Reweight()")
        self.handle.step(self._pars)

    def SynchroniseMemTransfer(self):
        self.handle.synchronize()
        if self._standalone:
            self.cpu_total_weights, _ = self.handle.read_event_weights()

    def retPointer(self, event):
        """View of the event's total spline weight (valid after SynchroniseMemTransfer)."""
        return self.cpu_total_weights[event:event + 1]

    def FindSplineSegment(self, pars=None):
        return self.handle.find_segments(self._pars if pars is None else pars)


class BinnedSplineHandler:
    """The binned SplineBase implementation (Splines/BinnedSplineHandler.h) on the B200, always attached
    to a :class:`SampleHandlerFD`: its evaluation runs inside ``Reweight()``; ``weightvec_Monolith`` is a
    lazy host mirror."""

    def __init__(self, spl, handle, f64=False):
        self.handle = handle
        self.n_params = int(spl["n_params"])
        handle.upload_binned_splines(spl, f64=f64)

    def GetName(self):
        return "BinnedSplineHandler"

    def GetNParams(self):
        return self.n_params

    def setSplinePointers(self, spline_pars_array):
        assert spline_pars_array.dtype == np.float64 and spline_pars_array.size == self.n_params
        self._pars = spline_pars_array

    @property
    def weightvec_Monolith(self):
        return self.handle.read_binned_weights()

    def FindSplineSegment(self, pars=None):
        return self.handle.find_segments(self._pars if pars is None else pars)


def slice_monolith(spl, e0, e1):
    """Events [e0,e1) of a reference-layout monolith, offsets rebased to the chunk."""
    cnt = np.asarray(spl["nParamPerEvent"]).reshape(-1, 2)[:, 0].astype(np.int64)
    cntl = np.asarray(spl["nParamPerEvent_tf1"]).reshape(-1, 2)[:, 0].astype(np.int64)
    sc = np.concatenate([[0], np.cumsum(cnt)])
    sl = np.concatenate([[0], np.cumsum(cntl)])
    a, b = int(sc[e0]), int(sc[e1])
    la, lb = int(sl[e0]), int(sl[e1])
    kn = np.asarray(spl["nKnots_arr"]).astype(np.uint64)
    total_knots = np.asarray(spl["coeff_many"]).size // 4
    k0 = int(kn[a]) if b > a else 0
    k1 = (int(kn[b]) if b < kn.size else total_knots) if b > a else 0
    npe = np.zeros(2 * (e1 - e0), np.uint32)
    npe[0::2] = cnt[e0:e1]
    npe[1::2] = (sc[e0:e1] - a).astype(np.uint32)
    npl = np.zeros(2 * (e1 - e0), np.uint32)
    npl[0::2] = cntl[e0:e1]
    npl[1::2] = (sl[e0:e1] - la).astype(np.uint32)
    return dict(n_events=e1 - e0, nParamPerEvent=npe, paramNo_arr=np.asarray(spl["paramNo_arr"])[a:b],
                nKnots_arr=(kn[a:b] - np.uint64(k0)), coeff_many=np.asarray(spl["coeff_many"])[4 * k0:4 * k1],
                nParamPerEvent_tf1=npl, paramNo_tf1=np.asarray(spl["paramNo_tf1"])[la:lb],
                coeff_tf1=np.asarray(spl["coeff_tf1"])[2 * la:2 * lb])


class SampleHandlerFD:
    """Far-detector sample handler whose Reweight/GetLikelihood run as one fused device pass."""

    def __init__(self, edges, test_statistic=_lib.POISSON, update_w2=False, device=0, tile_events=0,
                 keep_event_weights=False, fused_llh=True, keep_kinematics=False, batch_kernel=True):
        """keep_kinematics: needed for handle.update_kinematics (functional shifts applied on the host)."""
        flags = ((_lib.FLAG_KEEP_EVENT_WEIGHTS if keep_event_weights else 0) | (0 if fused_llh else _lib.FLAG_NO_FUSED_LLH)
                 | (_lib.FLAG_KEEP_KINEMATICS if keep_kinematics else 0) | (0 if batch_kernel else _lib.FLAG_NO_BATCH_KERNEL))
        self.handle = _lib.Handle(device=device, test_statistic=test_statistic, update_w2=update_w2,
                                  tile_events=tile_events, flags=flags)
        self.handle.upload_binning(edges)
        self.n_bins = self.handle.n_bins
        self.n_samples = self.handle.n_samples
        self.SplineHandler = None
        self._spline_pars = None
        self._norm_pars = None
        self._osc_w = None
        self._osc_dirty = False
        self._data = np.zeros(self.n_bins, np.float64)
        self._fused = fused_llh

    # -- wiring (what SampleHandlerFD::Initialise does, Samples/SampleHandlerFD.cpp:169-202) --------
    def SetupSplines(self, n_params, max_knots, coeff_x, n_pts, spl, chunk_events=None):
        self.SplineHandler = SMonolith(n_params, max_knots, coeff_x, n_pts, spl, handle=self.handle,
                                       chunk_events=chunk_events)
        return self.SplineHandler

    def SetupBinnedSplines(self, spl, f64=False):
        """f64: the reference's default build (M3::float_t = double); oscillation / static weights are then doubles."""
        self.SplineHandler = BinnedSplineHandler(spl, self.handle, f64=f64)
        self._f64 = bool(f64)
        return self.SplineHandler

    def SetBinnedSplinePointers(self, n_per_event, spline_index):
        """SampleHandlerFD::SetSplinePointers, binned arm (Samples/SampleHandlerFD.cpp:1196-1242); after SetupEvents."""
        self.handle.upload_event_binned_splines(n_per_event, spline_index)

    def SetupEvents(self, sample_id, kin, norm_idx=None, n_norm_per_event=0, norm_pars=None, osc_w=None,
                    osc_idx=None, static_w=None):
        """norm_pars / osc_w are the live arrays the reference's per-event pointers point into."""
        self._norm_pars = norm_pars
        self._osc_w = osc_w
        self.handle.upload_events(sample_id, kin, norm_idx, n_norm_per_event,
                                  0 if norm_pars is None else norm_pars.size, osc_w is not None, osc_idx,
                                  0 if osc_w is None else osc_w.size, static_w)
        self._osc_dirty = osc_w is not None

    def SetSelection(self, cuts, values=None):
        """StoredSelection (Samples/SampleHandlerFD.cpp:162): cuts = [(sample, var, lower, upper), ...];
        values[n_vars, n_events] = ReturnKinematicParameter(var, event) for the cut variables."""
        self.handle.upload_selection(cuts, values)

    def SetSplinePointers(self, spline_pars_array):
        self._spline_pars = spline_pars_array
        if self.SplineHandler is not None:
            self.SplineHandler.setSplinePointers(spline_pars_array)

    def OscillatorEvaluated(self):
        """Tell the handler the oscillation-weight array changed (Oscillator->Evaluate() ran)."""
        self._osc_dirty = True

    # -- reference API -------------------------------------------------------------------------
    def Reweight(self):
        if getattr(self, "_f64", False):
            if self._osc_dirty and self._osc_w is not None:
                self.handle.upload_osc_f64(self._osc_w)
            self.handle.step(self._spline_pars, self._norm_pars, None, mode="fused" if self._fused else "fill")
            self._osc_dirty = False
            return
        osc = self._osc_w if self._osc_dirty else None
        self.handle.step(self._spline_pars, self._norm_pars, osc, mode="fused" if self._fused else "fill")
        self._osc_dirty = False

    def GetLikelihood(self):
        return self.handle.llh()

    def GetSampleLikelihood(self, isample):
        return float(self.handle.llh(per_sample=True)[1][isample])

    def AddData(self, data):
        self._data = np.ascontiguousarray(data, np.float64).copy()
        self.handle.upload_data(self._data)

    def SetTestStatistic(self, ts):
        self.handle.set_test_statistic(ts)

    def GetMCArray(self):
        return self.handle.read_hist()[0]

    def GetW2Array(self):
        return self.handle.read_hist()[1]

    def GetDataArray(self):
        return self._data

    def GetNEvents(self):
        return self.handle.n_events

    def GetEventWeight(self):
        """Per-event (spline weight, total weight) of the last Reweight (lazy D2H)."""
        return self.handle.read_event_weights()

    def GetEventBins(self):
        return self.handle.read_event_bins()


def build_from_workload(w, e0=0, e1=None, with_osc=True, update_w2=False, test_statistic=None, device=0,
                        tile_events=0, keep_event_weights=False, chunk_events=None, fused_llh=True, batch_kernel=True):
    """B200 SampleHandlerFD + SMonolith wired on a synthetic workload (mirror of the oracle's helper)."""
    from . import synth
    e1 = w.n_events if e1 is None else e1
    typ, npts, cx = synth.param_layout(w)
    spl = synth.make_splines(w, e0, e1)
    ev = synth.make_events(w, e0, e1)
    sh = SampleHandlerFD(synth.bin_edges(w), w.test_statistic if test_statistic is None else test_statistic,
                         update_w2, device, tile_events, keep_event_weights, fused_llh, batch_kernel=batch_kernel)
    sh.SetupSplines(w.n_params, w.n_knots, cx, npts, spl, chunk_events=chunk_events)
    pars = np.zeros(w.n_params, np.float64)
    norm = np.ones(max(w.n_norm_params, 1), np.float64)[:w.n_norm_params]
    osc = synth.make_osc(w, 0, e0, e1) if with_osc else None
    sh.SetupEvents(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, norm if w.n_norm_params else None,
                   osc, None, ev["static_w"])
    sh.SetSplinePointers(pars)
    return sh, dict(typ=typ, npts=npts, coeff_x=cx, spl=spl, ev=ev, pars=pars, norm=norm, osc=osc)


def build_binned_from_workload(w, update_w2=True, test_statistic=None, device=0, keep_event_weights=False, fused_llh=True,
                               f64=False, spl=None, ev=None):
    """B200 SampleHandlerFD + BinnedSplineHandler wired on a synthetic binned-spline workload
    (f64: the reference's default build, M3::float_t = double).  spl / ev: use these arrays instead of generating them."""
    from .synth import binned as B
    spl = B.make_binned_splines(w, f64=f64) if spl is None else spl
    ev = B.make_binned_events(w, f64=f64) if ev is None else ev
    sh = SampleHandlerFD(B.bin_edges(w), w.test_statistic if test_statistic is None else test_statistic, update_w2, device,
                         0, keep_event_weights, fused_llh)
    sh.SetupBinnedSplines(spl, f64=f64)
    pars = np.zeros(w.n_systs, np.float64)
    norm = np.ones(w.n_norm_params, np.float64)
    osc = B.make_osc(w, 0, f64=f64)
    if f64:
        # the float entry points only fix the shapes here; the double arrays follow
        sh.handle.upload_events(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, norm.size, True, None, 0, None)
        sh._norm_pars, sh._osc_w, sh._osc_dirty = norm, osc, True
        sh.SetBinnedSplinePointers(ev["n_per_event"], ev["spline_index"])
        sh.handle.upload_event_weights_f64(ev["static_w"])
        sh.SetSplinePointers(pars)
        return sh, dict(spl=spl, ev=ev, pars=pars, norm=norm, osc=osc)
    sh.SetupEvents(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, norm, osc, None, ev["static_w"])
    sh.SetBinnedSplinePointers(ev["n_per_event"], ev["spline_index"])
    sh.SetSplinePointers(pars)
    return sh, dict(spl=spl, ev=ev, pars=pars, norm=norm, osc=osc)
