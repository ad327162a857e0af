"""klu_topsort (host side of the C ABI) against the oracle's restatement of
TopSortCompactLatticeIfNeeded [ext] (oracle/klu_oracle.cc OpenFstTopOrder, a recursive depth-first
visit) and against the Python twin in kaldi-lattice-utils_b200/lattice.py."""
import ctypes as C

import numpy as np
import pytest


def _call(klu, lat):
    lib = klu.binding.lib()
    a = {k: np.ascontiguousarray(getattr(lat, k)).copy() for k in
         ("src", "dst", "label", "dur", "graph", "acoustic", "fin_graph", "fin_acoustic", "fin_dur")}
    order = np.zeros(lat.nstates, np.int32)
    p = lambda x: C.c_void_p(x.ctypes.data)
    rc = lib.klu_topsort(C.c_int32(lat.nstates), C.c_int64(lat.narcs), p(a["src"]), p(a["dst"]), p(a["label"]),
                         p(a["dur"]), p(a["graph"]), p(a["acoustic"]), p(a["fin_graph"]), p(a["fin_acoustic"]),
                         p(a["fin_dur"]), p(order))
    return rc, a, order


def _random_dag(klu, rng, n, shuffle=True):
    arcs = []
    for s in range(n - 1):
        for _ in range(rng.randint(1, 4)):
            d = rng.randint(s + 1, min(n, s + 4))
            arcs.append((s, d, int(rng.randint(0, 6)), float(rng.uniform(0, 3)), float(rng.uniform(0, 3)), d - s))
    perm = np.arange(n)
    if shuffle:
        perm[1:] = rng.permutation(np.arange(1, n))  # the start state stays 0
    arcs = [(int(perm[s]), int(perm[d]), w, g, a, t) for (s, d, w, g, a, t) in arcs]
    arcs.sort(key=lambda x: x[0])  # grouped by source, as OpenFst stores them
    return klu.make_lattice("dag", n, arcs, {int(perm[n - 1]): (0.5, 0.25)})


def _apply_order(klu, lat, order):
    """The lattice renumbered by order[old] = new, arcs regrouped by new source (stable)."""
    arcs = [(order[s], order[d], int(w), float(g), float(a), int(t)) for s, d, w, g, a, t in
            zip(lat.src.tolist(), lat.dst.tolist(), lat.label.tolist(), lat.graph.tolist(), lat.acoustic.tolist(),
                lat.dur.tolist())]
    arcs.sort(key=lambda x: x[0])
    finals = {order[s]: (float(lat.fin_graph[s]), float(lat.fin_acoustic[s])) for s in range(lat.nstates)
              if not np.isinf(lat.fin_graph[s])}
    return klu.make_lattice(lat.key, lat.nstates, arcs, finals)


@pytest.mark.parametrize("seed", range(12))
def test_topsort_matches_oracle(klu, ora, seed):
    """State ids after the sort decide what lattice-prune-dyn-beam writes: klu_topsort must
    number the states exactly as the oracle's fst::TopSort restatement does."""
    rng = np.random.RandomState(100 + seed)
    lat = _random_dag(klu, rng, 3 + seed * 5)
    want_order = ora.top_order(lat)
    rc, got, order = _call(klu, lat)
    assert rc == 0
    assert order.tolist() == want_order
    want = _apply_order(klu, lat, want_order)
    for k in ("src", "dst", "label", "dur"):
        assert np.array_equal(got[k], np.asarray(getattr(want, k))), k
    for k in ("graph", "acoustic", "fin_graph", "fin_acoustic"):
        assert np.array_equal(got[k], np.asarray(getattr(want, k), dtype=np.float32)), k


@pytest.mark.parametrize("seed", range(6))
def test_topsort_matches_python(klu, seed):
    rng = np.random.RandomState(seed)
    lat = _random_dag(klu, rng, 3 + seed * 7)
    want = klu.topsort(lat)
    rc, got, order = _call(klu, lat)
    assert rc == 0
    assert (got["src"] < got["dst"]).all() and (np.diff(got["src"]) >= 0).all()
    for k in ("src", "dst", "label", "dur", "graph", "acoustic", "fin_graph", "fin_acoustic", "fin_dur"):
        assert np.array_equal(got[k], np.asarray(getattr(want, k))), k
    assert sorted(order.tolist()) == list(range(lat.nstates))


def test_topsort_identity_when_sorted(klu):
    lat = _random_dag(klu, np.random.RandomState(3), 12, shuffle=False)
    rc, got, order = _call(klu, lat)
    assert rc == 0 and order.tolist() == list(range(12))
    assert np.array_equal(got["src"], lat.src) and np.array_equal(got["graph"], lat.graph)


def test_topsort_cycle_is_an_error(klu):
    lat = klu.make_lattice("cyc", 3, [(0, 1, 1, 0.0, 0.0, 1), (1, 2, 2, 0.0, 0.0, 1), (2, 1, 3, 0.0, 0.0, 1)],
                           {2: (0.0, 0.0)})
    rc, _, _ = _call(klu, lat)
    assert rc != 0
    assert b"cyclic" in klu.binding.lib().klu_last_error()
