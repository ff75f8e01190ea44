"""Generates tests/golden/ref_host_binning.npz by running the REFERENCE's own binning code
(Samples/SampleStructs.h compiled from /root/reference into oracle/_ref/libm3ref_host.so by oracle/ref_host/Makefile).
Run here (the container that has /root/reference):   python tests/golden/make_ref_host_golden.py
The vectors travel; the oracle is checked against them on any box (tests/test_reference_host.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_host_binding as RH  # noqa: E402


def cases():
    rng = np.random.default_rng(2024)
    c = {}
    c["uni1d"] = [np.array([0.0, 0.1, 0.25, 0.6, 1.0, 1.7, 2.5, 4.0, 10.0])]
    c["uni2d"] = [np.concatenate([[0.0], np.sort(rng.uniform(0.05, 2.95, 14)), [3.0]]), np.linspace(0.0, np.pi, 8)]
    c["uni3d"] = [np.array([-1.0, 0.0, 0.5, 2.0]), np.array([0.0, 1.0, 3.0]), np.array([10.0, 20.0, 25.0, 70.0, 71.0])]

    def boxes(nx, x_max, y_max):
        xs = np.concatenate([[0.0], np.sort(rng.uniform(0.1, x_max - 0.1, nx - 1)), [x_max]])
        out = []
        for i in range(nx):
            ny = int(rng.integers(1, 7))
            ys = np.concatenate([[0.0], np.sort(rng.uniform(0.05, y_max - 0.05, ny - 1)), [y_max]])
            for j in range(ny):
                out.append([[xs[i], xs[i + 1]], [ys[j], ys[j + 1]]])
        return np.array(out)
    c["box_a"] = boxes(7, 3.0, np.pi)
    c["box_b"] = boxes(23, 10.0, 1.0)
    c["box_doc"] = np.array([[[0, 1], [0, 1]], [[1, 3], [0, 0.5]], [[1, 3], [0.5, 1]], [[0, 1], [1, 2]], [[1, 3], [1, 2]]], float)
    return c


def points(spec, rng, n=4000):
    """random points + every edge value + values one ulp around edges; nominal bins: right, off by one, far, -1"""
    if isinstance(spec, np.ndarray):
        nd = spec.shape[1]
        lo, hi = spec[:, :, 0].min(0), spec[:, :, 1].max(0)
        special = [np.unique(spec[:, d, :]) for d in range(nd)]
    else:
        nd = len(spec)
        lo, hi = np.array([e[0] for e in spec]), np.array([e[-1] for e in spec])
        special = [np.asarray(e) for e in spec]
    kin = np.empty((nd, n))
    for d in range(nd):
        span = hi[d] - lo[d]
        kin[d] = rng.uniform(lo[d] - 0.1 * span, hi[d] + 0.1 * span, n)
        sp = np.concatenate([special[d], np.nextafter(special[d], -np.inf), np.nextafter(special[d], np.inf)])
        idx = rng.choice(n, size=min(n // 2, 6 * sp.size), replace=False)
        kin[d, idx] = rng.choice(sp, idx.size)
    return kin


def main():
    rng = np.random.default_rng(7)
    out = {}
    for name, spec in cases().items():
        ref = RH.RefBinning(spec)
        kin = points(spec, rng)
        nd = kin.shape[0]
        edges = [ref.axis_edges(d) for d in range(nd)]
        nb = [e.size - 1 for e in edges]
        # nominal bins: the true one (upper_bound-1, or -1 out of range), then perturbed
        true_nom = np.empty((nd, kin.shape[1]), np.int32)
        for d in range(nd):
            t = np.searchsorted(edges[d], kin[d], side="right") - 1
            t[(kin[d] < edges[d][0]) | (kin[d] >= edges[d][-1])] = -1
            true_nom[d] = t
        pert = true_nom + rng.integers(-2, 3, true_nom.shape).astype(np.int32)
        for d in range(nd):
            pert[d] = np.clip(pert[d], -1, nb[d] - 1)
        out[f"{name}/spec"] = spec if isinstance(spec, np.ndarray) else np.array([0.0])
        if not isinstance(spec, np.ndarray):
            for d, e in enumerate(spec):
                out[f"{name}/in_edges{d}"] = np.asarray(e, float)
        out[f"{name}/kin"] = kin
        out[f"{name}/nom_true"] = true_nom
        out[f"{name}/nom_pert"] = pert
        for d in range(nd):
            out[f"{name}/edges{d}"] = edges[d]
            out[f"{name}/findbin_true{d}"] = ref.find_bin(d, kin[d], true_nom[d])
            out[f"{name}/findbin_pert{d}"] = ref.find_bin(d, kin[d], pert[d])
        out[f"{name}/bin_true"] = ref.find_sample_bin(kin, true_nom)
        out[f"{name}/bin_pert"] = ref.find_sample_bin(kin, pert)
        out[f"{name}/n_bins"] = np.array([ref.n_bins])
        if isinstance(spec, np.ndarray):
            n_mega = int(np.prod(nb))
            gm = [ref.grid_mapping(g) for g in range(n_mega)]
            out[f"{name}/grid_len"] = np.array([len(x) for x in gm], np.int32)
            out[f"{name}/grid_idx"] = np.array([b for x in gm for b in x], np.int32)
        ref.close()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_host_binning.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in list(out.items())[:6]})


if __name__ == "__main__":
    main()
