"""The fitter's call surface: a Metropolis chain written the way MR2T2 drives its samples
(Fitters/MR2T2.cpp:25-116: ProposeStep -> samples[i]->Reweight() -> samples[i]->GetLikelihood(), accept with
min(1, exp(logLCurr - logLProp)), rejected steps leave the parameters where they were).  Run once on the CPU oracle
and once on the B200 handlers with the same random stream: the accept/reject sequence must be identical and the
-lnL trace must agree to the north_star tolerance; since rejected steps do not move the parameters but DO move the
cached spline segments (SplineBase.cpp:76), the chain also exercises the history dependence of the segments."""
import numpy as np
import pytest

from mach3_b200 import handlers, synth
from oracle import binding as O

pytestmark = pytest.mark.gpu


class _Params:                      # ParameterHandler stand-in: flat proposal, reflecting at the spline range
    def __init__(self, w, seed):
        self.rng = np.random.default_rng(seed)
        self.cur_sp, self.cur_nm = synth.proposal(w, -1)
        self.prop_sp, self.prop_nm = self.cur_sp.copy(), self.cur_nm.copy()

    def ProposeStep(self):
        self.prop_sp = np.clip(self.cur_sp + self.rng.normal(0, 0.15, self.cur_sp.size), -2.9, 2.9)
        if self.rng.random() < 0.1:          # now and then land exactly on a knot
            self.prop_sp[self.rng.integers(self.cur_sp.size)] = float(self.rng.integers(-2, 3))
        self.prop_nm = np.clip(self.cur_nm + self.rng.normal(0, 0.02, self.cur_nm.size), 0.5, 1.5)

    def GetLikelihood(self):                 # Gaussian prior, -lnL
        return 0.5 * float(np.sum(self.prop_sp ** 2)) + 0.5 * float(np.sum(((self.prop_nm - 1) / 0.1) ** 2))

    def AcceptStep(self):
        self.cur_sp, self.cur_nm = self.prop_sp.copy(), self.prop_nm.copy()


def _chain(sample, set_pars, n_steps, w, seed):
    pars = _Params(w, seed)
    set_pars(pars.cur_sp, pars.cur_nm)
    sample.Reweight()
    logLCurr = pars.GetLikelihood() + sample.GetLikelihood()
    trace, accepted = [], []
    for _ in range(n_steps):
        pars.ProposeStep()
        set_pars(pars.prop_sp, pars.prop_nm)
        sample.Reweight()                                # MR2T2::ProposeStep
        logLProp = pars.GetLikelihood() + sample.GetLikelihood()
        acc = min(1.0, np.exp(logLCurr - logLProp))      # MR2T2::AcceptanceProbability
        ok = pars.rng.random() < acc
        if ok:
            pars.AcceptStep(); logLCurr = logLProp
        trace.append(logLProp); accepted.append(ok)
    return np.array(trace), np.array(accepted)


@pytest.mark.parametrize("wl,n", [("CFG1", 20_000), ("SPARSE_RUNS", 20_000)])
def test_metropolis_chain_is_identical_on_oracle_and_b200(wl, n):
    O.set_multithread(False)
    w = getattr(synth, wl).scaled(n)
    mono, osh, od = O.build_from_workload(w)
    gsh, gd = handlers.build_from_workload(w)

    def set_o(sp, nm):
        mono.set_params(sp); osh.norm_vals[:] = nm

    def set_g(sp, nm):
        gd["pars"][:] = sp; gd["norm"][:] = nm

    set_o(*synth.proposal(w, -1)); osh.Reweight()
    data = np.random.default_rng(21).poisson(osh.mc).astype(np.float64)
    osh.AddData(data); gsh.AddData(data)
    t_o, a_o = _chain(osh, set_o, 150, w, seed=5)
    t_g, a_g = _chain(gsh, set_g, 150, w, seed=5)
    np.testing.assert_array_equal(a_g, a_o)              # same decisions at every step
    np.testing.assert_allclose(t_g, t_o, rtol=1e-9, atol=1e-9)
    assert 0.05 < a_o.mean() < 0.99
    O.set_multithread(True)
