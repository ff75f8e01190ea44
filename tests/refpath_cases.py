"""Seeded inputs for the reference-host-path parity tests (tests/test_reference_path.py) and the script that makes
their golden vectors (tests/golden/make_ref_host_path.py).  Per-event response functions in the shape MaCh3 hands to
SMonolith's constructor (std::vector<std::vector<TResponseFunction_red*>>, Splines/SplineMonolith.cpp:36-49):
TSpline3_red knots {x, y, b, c, d} and TF1_red {a, b}.  numpy's PCG64 streams are stable across versions."""
import numpy as np


def _knots(p, k, representable):
    if representable:                     # multiples of 1/4: exact in float, the usual sigma grids
        return (np.arange(k) - (k - 1) / 2.0) * (0.5 if k > 3 else 1.0)
    lo, hi = -1.3 - 0.1 * p, 1.7 + 0.07 * p
    return np.linspace(lo, hi, k)         # plain decimals: NOT exact in float


def make_case(name):
    """-> dict(type[P], npts[E,P], vals[sum,5], pars[T,P])"""
    if name == "mixed":
        E, P, n_tf1, cover, seed, flat = 600, 14, 4, 0.6, 11, False
        K = [2, 3, 5, 7, 9, 4, 6, 8, 3, 5]
    elif name == "flat":
        E, P, n_tf1, cover, seed, flat = 400, 9, 2, 0.5, 12, True
        K = [5, 3, 7, 4, 6, 2, 9]
    elif name == "dense2":
        E, P, n_tf1, cover, seed, flat = 300, 3, 0, 1.0, 13, False
        K = [2, 2, 2]
    else:
        raise KeyError(name)
    rng = np.random.default_rng(seed)
    n_spl = P - n_tf1
    typ = np.array([0] * n_spl + [1] * n_tf1, np.int32)
    x_of = [_knots(p, K[p], representable=(p % 2 == 0)) for p in range(n_spl)]
    dead = 3 if flat else -1                                   # a parameter no event responds to
    npts = np.zeros((E, P), np.int32)
    rows = []
    for e in range(E):
        for p in range(P):
            has = (e == 0 or rng.random() < cover) and p != dead
            if flat and e > 0 and e % 7 == 0:
                has = False                                    # events without any response
            if not has:
                continue
            if typ[p]:
                npts[e, p] = 2
                a, b = rng.normal(0.0, 0.05), 1.0 + rng.normal(0.0, 0.02)
                rows.append(np.array([[0, a, 0, 0, 0], [0, b, 0, 0, 0]], np.float64))
            elif flat and e > 0 and rng.random() < 0.15:
                npts[e, p] = 1                                 # a one-knot ("flat") spline: skipped at :151
                rows.append(np.array([[x_of[p][0], 1.0, 0, 0, 0]], np.float64))
            else:
                k = K[p]
                npts[e, p] = k
                r = np.zeros((k, 5))
                r[:, 0] = x_of[p]
                r[:, 1] = 1.0 + 0.15 * rng.normal(size=k)
                r[:, 2] = 0.10 * rng.normal(size=k)
                r[:, 3] = 0.05 * rng.normal(size=k)
                r[:, 4] = 0.02 * rng.normal(size=k)
                rows.append(r)
    vals = np.concatenate(rows, 0)

    # the proposals: a random walk (exercises the cached-segment shortcut) interleaved with boundary cases
    T = 40
    pars = np.zeros((T, P))
    walk = np.zeros(P)
    for t in range(1, T):
        walk = walk + rng.normal(0.0, 0.6, P)
        walk = np.clip(walk, -4.5, 4.5)
        pars[t] = walk
    def knot(p, k):
        return x_of[p][min(k, len(x_of[p]) - 1)] if p < n_spl else 0.3
    for p in range(P):
        pars[5, p] = knot(p, 0)                                             # exactly the first knot
        pars[9, p] = knot(p, 99)                                            # exactly the last knot
        pars[13, p] = knot(p, 1)                                            # an interior knot (as a double)
        pars[17, p] = float(np.float32(knot(p, 1)))                         # ... its float rounding
        pars[21, p] = float(np.nextafter(np.float32(knot(p, 1)), np.float32(9)))    # one float ulp above
        pars[25, p] = float(np.nextafter(np.float32(knot(p, 1)), np.float32(-9)))   # one float ulp below
        pars[29, p] = -10.0 if p % 2 else 10.0                              # far outside the knots
        pars[33, p] = knot(p, 2) + 1e-12                                    # a double hair above a knot
    return dict(type=typ, npts=npts, vals=vals, pars=pars)


CASES = ("mixed", "flat", "dense2")


def stat_inputs():
    """(data, mc, w2) triples for SampleHandlerBase::GetTestStatLLH: the branches around data == 0, mc == 0,
    M3::_LOW_MC_BOUND_ (1e-5), w2 == 0, and ordinary bins.  128 = 2 x 64 so that the device test can give every
    triple its own one-bin sample."""
    rng = np.random.default_rng(5)
    special = [(d, m) for d in (0.0, 1e-6, 1e-5, 2e-5, 1.0, 7.0, 250.0) for m in (0.0, 1e-7, 1e-5, 3e-5, 0.5, 7.0, 260.0)]
    ns, nr = len(special), 128 - len(special)
    data = np.array([s[0] for s in special] + list(rng.poisson(rng.uniform(0.1, 40.0, nr)).astype(float)))
    mc = np.array([s[1] for s in special] + list(rng.uniform(0.05, 45.0, nr)))
    w2 = np.concatenate([np.where(np.arange(ns) % 3 == 0, 0.0, mc[:ns] * rng.uniform(0.001, 0.5, ns)),
                         mc[ns:] * rng.uniform(0.001, 0.2, nr)])
    return data, mc, w2


# ---- the full per-step path: SampleHandlerFD::Reweight + GetLikelihood over the "mixed" monolith --------------------
FD_STEPS = 12
N_NORM = 6
NPE = 3


def fd_edges():
    """Three samples: 1-D uniform, 2-D uniform, 2-D non-uniform boxes (with a gap and an overlap-free L shape)."""
    boxes = np.array([[[0.0, 1.0], [0.0, 1.0]], [[1.0, 2.5], [0.0, 0.5]], [[1.0, 2.5], [0.5, 1.0]],
                      [[0.0, 0.7], [1.0, 2.0]], [[0.7, 2.5], [1.0, 2.0]], [[2.7, 3.0], [0.0, 2.0]]])
    return [[np.linspace(0.0, 3.0, 13)], [np.array([0.0, 0.4, 0.9, 1.5, 2.2, 3.0]), np.linspace(0.0, 2.0, 5)], boxes]


def fd_case():
    """Events for make_case("mixed"): sample, kinematics (some outside every bin, some on edges), up to NPE norm
    pointers, an oscillation weight and a static weight per event (a few zero / negative: the `<= 0` skip),
    FD_STEPS proposals, and one set of shifted kinematics (what functional parameters would write)."""
    c = make_case("mixed")
    E = c["npts"].shape[0]
    rng = np.random.default_rng(21)
    sample_id = rng.integers(0, 3, E).astype(np.int32)
    kin = np.zeros((2, E))
    kin[0] = rng.uniform(-0.2, 3.2, E)
    kin[1] = rng.uniform(-0.1, 2.1, E)
    on_edge = rng.integers(0, E, 40)
    kin[0, on_edge] = np.linspace(0.0, 3.0, 13)[rng.integers(0, 13, 40)]
    norm_idx = rng.integers(-1, N_NORM, (E, NPE)).astype(np.int16)
    static_w = rng.uniform(0.5, 1.5, E).astype(np.float32)
    static_w[rng.integers(0, E, 9)] = 0.0
    static_w[rng.integers(0, E, 5)] = -0.3
    osc = rng.random((FD_STEPS, E)).astype(np.float32)
    norm = np.clip(rng.normal(1.0, 0.1, (FD_STEPS, N_NORM)), 0.5, 1.5)
    kin_shift = kin + rng.normal(0.0, 0.08, kin.shape)
    return dict(mono=c, sample_id=sample_id, kin=kin, kin_shift=kin_shift, norm_idx=norm_idx, static_w=static_w,
                osc=osc, norm=norm, pars=c["pars"][[0, 1, 2, 5, 9, 13, 17, 21, 25, 29, 33, 38]])


def binned_workload():
    from mach3_b200.synth import binned as B
    return B.CFG4_SMALL.scaled(n_events=4000, n_grid=150)


BINNED_STEPS = (-1, 0, 1, 2, -2, -3, -4, 3)


# ---- selection cuts: SampleHandlerFD::IsEventSelected (Samples/SampleHandlerFD.cpp:281-294) over fd_case() -----------
SEL_STEPS = 6
SEL_SHIFT_AT = 4


def selection_case():
    """StoredSelection for the three samples of fd_edges() and the kinematic table the cuts read.  kin4[v, e]: rows 0/1
    are the binning variables of fd_case(), rows 2/3 exist only to be cut on.  Cuts (sample, variable, lower, upper):
    sample 0 on its binning variable and on variable 2; sample 1 on variable 3, which holds values EXACTLY on both
    bounds (lower passes, upper fails); sample 2 on its second binning variable and variable 2.  kin4_shift: every row
    moved (what functional parameters write before IsEventSelected runs, :359-361)."""
    f = fd_case()
    E = f["sample_id"].size
    rng = np.random.default_rng(41)
    kin4 = np.zeros((4, E))
    kin4[:2] = f["kin"]
    kin4[2] = rng.uniform(0.0, 2.0, E)
    kin4[3] = rng.uniform(0.0, 1.0, E)
    on_lo, on_hi = rng.integers(0, E, 25), rng.integers(0, E, 25)
    kin4[3, on_lo] = 0.25
    kin4[3, on_hi] = 0.75
    kin4[3, rng.integers(0, E, 3)] = np.nan                  # NaN fails neither comparison: the reference selects it
    cuts = [(0, 0, 0.5, 2.4), (0, 2, 0.1, 1.8), (1, 3, 0.25, 0.75), (2, 1, 0.2, 1.9), (2, 2, 0.3, 2.0)]
    kin4_shift = kin4.copy()
    kin4_shift[:2] = f["kin_shift"]
    kin4_shift[2] = kin4[2] + rng.normal(0.0, 0.1, E)
    kin4_shift[3] = np.where(np.isnan(kin4[3]), np.nan, np.clip(kin4[3] + rng.normal(0.0, 0.05, E), 0.0, 1.0))
    return dict(kin4=kin4, kin4_shift=kin4_shift, cuts=cuts)


# ---- functional ("shift") parameters: SampleHandlerFD::ApplyShifts (Samples/SampleHandlerFD.cpp:545-564) ---------------
SHIFT_STEPS = 6
N_SHIFT = 3


def shift_case():
    """Three linear functional parameters on the events of selection_case(): parameter 0 scales the first binning
    variable (coef = the variable itself: an energy-scale shift), parameter 1 adds a per-event offset to the second
    binning variable for ~60 % of the events, parameter 2 moves cut-only variable 2 (so events cross the selection
    bounds).  coef[s, e] NaN = the event is not in the parameter's funcParsGrid list.  values[t, s]: the parameter values
    per step (step 0: all zero = nominal)."""
    sel = selection_case()
    E = sel["kin4"].shape[1]
    rng = np.random.default_rng(51)
    coef = np.full((N_SHIFT, E), np.nan)
    coef[0] = sel["kin4"][0]
    on1 = rng.random(E) < 0.6
    coef[1, on1] = rng.normal(0.0, 0.5, int(on1.sum()))
    on2 = rng.random(E) < 0.5
    coef[2, on2] = rng.uniform(0.2, 1.0, int(on2.sum()))
    target = np.array([0, 1, 2], np.int32)
    values = rng.normal(0.0, 0.08, (SHIFT_STEPS, N_SHIFT))
    values[0] = 0.0
    values[3, 2] = 0.9                      # a large move of the cut variable
    return dict(target=target, coef=coef, values=values)


def shift_entries(sh, n_kin_rows=2, mirror_cut_rows=True):
    """The coefficient matrix of shift_case() as per-event entry lists (the layout of m3b_upload_linear_shifts and of the
    oracle): n_per_event[E], shift_par[], target[], coef[] in parameter order.  Kinematic column c of the reference
    harness is binning row c (c < n_kin_rows) AND row c of the 4-row cut-variable table the selection tests use, so a
    shift of a binning variable is entered twice (target c and target n_kin_rows + c) when mirror_cut_rows."""
    n_pars, E = sh["coef"].shape
    npe, par, tgt, cf = np.zeros(E, np.uint32), [], [], []
    for e in range(E):
        for s in range(n_pars):
            c = sh["coef"][s, e]
            if np.isnan(c):
                continue
            col = int(sh["target"][s])
            targets = [col, n_kin_rows + col] if (col < n_kin_rows and mirror_cut_rows) else [col if col < n_kin_rows else n_kin_rows + col]
            for t in targets:
                par.append(s); tgt.append(t); cf.append(c); npe[e] += 1
    return npe, np.array(par, np.int32), np.array(tgt, np.int32), np.array(cf, np.float64)
