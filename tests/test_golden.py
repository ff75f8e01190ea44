"""Oracle vs golden vectors produced by the reference's own CUDA kernels on a B200
(tests/golden/ref_gpu_weights.npz, made by tests/golden/make_ref_gpu_golden.py)."""
import os

import numpy as np
import pytest

from mach3_b200 import synth
from oracle import binding as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_gpu_weights.npz")


def _cases():
    if not os.path.exists(GOLD):
        return []
    z = np.load(GOLD)
    return sorted({k.split(".")[0] for k in z.files})


@pytest.mark.skipif(not os.path.exists(GOLD), reason="golden vectors not generated yet")
@pytest.mark.parametrize("name", _cases())
@pytest.mark.parametrize("build", ["serial", "multithread"])
def test_oracle_matches_reference_gpu_golden(name, build):
    from tests.golden.make_ref_gpu_golden import inputs_digest
    z = np.load(GOLD)
    w = getattr(synth, name).scaled(int(z[f"{name}.n_events"]))
    typ, npts, cx = synth.param_layout(w)
    spl = synth.make_splines(w)
    assert inputs_digest(spl, cx) == bytes(z[f"{name}.digest"]).decode(), "synthetic generator drifted from the golden inputs"
    O.set_multithread(build == "multithread")
    try:
        mono = O.SMonolith(w.n_params, w.n_knots, cx, npts, spl)
        for i, st in enumerate(z[f"{name}.steps"]):
            mono.set_params(z[f"{name}.pars.{i}"])
            mono.Evaluate()
            np.testing.assert_array_equal(mono.segments, z[f"{name}.segments.{i}"])
            np.testing.assert_array_equal(mono.param_values, z[f"{name}.values.{i}"])
            gold = z[f"{name}.weights.{i}"]
            if build == "serial":
                np.testing.assert_array_equal(mono.total_weights, gold)          # bit-exact
            else:
                np.testing.assert_allclose(mono.total_weights, gold, rtol=1e-5, atol=0)
    finally:
        O.set_multithread(True)
