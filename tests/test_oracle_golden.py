"""Pins the CPU oracle (oracle/klu_oracle.cc) to every golden vector the
reference holds for the hot path: the worked examples of kwsbin2/README.md, and
to brute-force path enumeration for the tools the reference has no golden for."""
import os

import numpy as np
import pytest

from util import GOLD, TOL, assert_rows_match, goldens


@pytest.fixture(scope="module")
def word_lat(klu):
    return klu.read_text_ark(os.path.join(GOLD, "lattice.ark.txt"))[0]


@pytest.fixture(scope="module")
def char_lat(klu):
    return klu.read_text_ark(os.path.join(GOLD, "lattice.char.ark.txt"))[0]


def test_readme_utterance_exact_text(klu, ora, word_lat):
    assert klu.format_tuples("lat1", ora.utterance(word_lat)).strip() == goldens()["utterance"]


def test_readme_segment_exact_text(klu, ora, word_lat):
    assert klu.format_tuples("lat1", ora.segment(word_lat)).strip() == goldens()["segment"]


def test_readme_position_exact_text(klu, ora, word_lat):
    assert klu.format_tuples("lat1", ora.position(word_lat)).strip() == goldens()["position"]


def _parse_char_golden(s):
    body = s.split(" ", 1)[1]
    rows = []
    for t in body.split(";"):
        f = t.split()
        rows.append((f[0], int(f[1]), int(f[2]), int(f[3]), float(f[4])))
    return rows


def test_readme_char_position(ora, char_lat):
    # keys, positions, segments and order exact; values within 1e-4 (the
    # reference computes this tool in float32 log semiring with determinize
    # delta, its README deviates from exact math by up to 5.9e-5)
    want = _parse_char_golden(goldens()["char_position"])
    got = ora.char_position(char_lat, [28])
    assert [r[:4] for r in got] == [r[:4] for r in want]
    for g, w in zip(got, want):
        assert abs(g[4] - w[4]) <= TOL


def test_readme_char_segment(ora, char_lat):
    # kwsbin2/README.md:182 (lattice-char-index-segment): strings,
    # segments and order exact; values within the reference's float32 noise
    body = goldens()["char_segment"].split(" ", 1)[1]
    want = []
    for t in body.split(";"):
        f = t.split()
        want.append((f[0], int(f[1]), int(f[2]), float(f[3])))
    got = ora.char_segment(char_lat, [28])
    assert [r[:3] for r in got] == [r[:3] for r in want]
    for g, w in zip(got, want):
        assert abs(g[3] - w[3]) <= TOL


def test_readme_state_times(klu, ora, word_lat):
    # kwsbin2/README.md:61-64; times come out of the segment keys
    times = goldens()["state_times"]
    seg = ora.segment(word_lat)
    t0s = {t0 for _, t0, _, _ in seg} | {t1 for _, _, t1, _ in seg}
    assert t0s <= set(times)


@pytest.mark.parametrize("seed", range(6))
def test_oracle_vs_bruteforce(klu, ora, seed):
    batch = klu.synth_batch("tiny", 4, seed=100 + seed)
    for lat in batch.lattices():
        seg = {(w, a, b): v for w, a, b, v in ora.segment(lat)}
        bs = ora.brute(ora.BRUTE_SEGMENT, lat)
        assert set(seg) == set(bs)
        assert max(abs(seg[k] - bs[k]) for k in seg) < 1e-9
        pos = {(w, p, 0): v for w, p, _, _, v in ora.position(lat)}
        bp = ora.brute(ora.BRUTE_POSITION, lat)
        assert set(pos) == set(bp)
        assert max(abs(pos[k] - bp[k]) for k in pos) < 1e-9
        utt = {(w, 0, 0): v for w, v in ora.utterance(lat)}
        bu = ora.brute(ora.BRUTE_UTTERANCE, lat)
        assert set(utt) == set(bu)
        # ComputeCompactLatticeBetas adds the two float weights in float
        assert max(abs(utt[k] - bu[k]) for k in utt) < 1e-4
        fr = ora.frame_post(lat)
        bf = ora.brute(ora.BRUTE_FRAME, lat)
        got = {(k, w, 0): p for k, row in enumerate(fr) for w, p in row}
        assert set(got) == set(bf)
        assert max(abs(got[k] - bf[k]) for k in got) < 1e-4


@pytest.mark.parametrize("seed", range(4))
def test_oracle_position_post_and_length_dist_vs_bruteforce(klu, ora, seed):
    """The SURVEY 8f tools have no golden in the reference: their restatements are checked
    against the all-paths enumeration.  Per (position, word) the posterior is the position
    tool's; the mass of transcripts with at least p words is the sum over the words at
    position p, and P(length = n) is the difference of two consecutive such masses."""
    import math
    batch = klu.synth_batch("tiny", 3, seed=300 + seed)
    for lat in batch.lattices():
        bp = ora.brute(ora.BRUTE_POSITION, lat)
        pp = ora.position_post(lat)
        got = {(w, k + 1): v for k, row in enumerate(pp) for w, v in row}
        assert set(got) == {(w, p) for w, p, _ in bp}
        # float32 rows, costs added in float (latbin/lattice-to-word-position-post.cc:104)
        assert max(abs(got[(w, p)] - bp[(w, p, 0)]) for w, p, _ in bp) < 1e-4
        for row in pp:  # each position's rows ordered by posterior
            assert all(a[1] >= b[1] for a, b in zip(row[:-1], row[1:]))
        at_least = {}
        for (w, p, _), v in bp.items():
            at_least[p] = at_least.get(p, 0.0) + math.exp(v)
        ld = dict(ora.length_dist(lat))
        for n, lp in ld.items():
            want = (at_least.get(n, 0.0) if n > 0 else 1.0) - at_least.get(n + 1, 0.0)
            assert abs(math.exp(lp) - want) < 1e-5, (n, lp, want)
        assert abs(sum(math.exp(v) for v in ld.values()) - 1.0) < 1e-5


def test_oracle_flags_consistency(klu, ora):
    # --acoustic-scale / --graph-scale / --insertion-penalty: equal to running the
    # default tool on a lattice whose float weights were transformed the way
    # ScaleLattice / AddWordInsPenToCompactLattice [ext] do it
    batch = klu.synth_batch("tiny", 3, seed=7)
    for lat in batch.lattices():
        g2 = (np.float64(0.7) * lat.graph.astype(np.float64)).astype(np.float32)
        a2 = (np.float64(np.float32(0.1)) * lat.acoustic.astype(np.float64)).astype(np.float32)
        g2 = np.where(lat.label != 0, g2 + np.float32(0.5), g2).astype(np.float32)
        fg = np.where(np.isinf(lat.fin_graph), lat.fin_graph,
                      (np.float64(np.float32(0.7)) * lat.fin_graph.astype(np.float64)).astype(np.float32))
        fa = np.where(np.isinf(lat.fin_acoustic), lat.fin_acoustic,
                      (np.float64(np.float32(0.1)) * lat.fin_acoustic.astype(np.float64)).astype(np.float32))
        g2 = (np.float64(np.float32(0.7)) * lat.graph.astype(np.float64)).astype(np.float32)
        g2 = np.where(lat.label != 0, g2 + np.float32(0.5), g2).astype(np.float32)
        lat2 = klu.Lattice(lat.key, lat.nstates, lat.src, lat.dst, lat.label, lat.dur, g2, a2,
                           fg.astype(np.float32), fa.astype(np.float32), lat.fin_dur)
        a = ora.segment(lat, acoustic_scale=0.1, graph_scale=0.7, insertion_penalty=0.5)
        b = ora.segment(lat2)
        assert a == b


def test_oracle_best_path2_readme_lattice(ora, word_lat):
    labels, cost = ora.best_path2(word_lat)
    # the 0.8-probability path "the dog is the man's best friend"
    assert labels == [2, 3, 5, 2, 6, 7, 8]
    assert abs(cost - 0.4) < 1e-6  # (1-0.8) + (1-0.8), float accumulation


def test_oracle_prune_dyn_beam_noop_and_prune(klu, ora):
    lat = klu.synth_batch("small", 1, seed=3)[0]
    r = ora.prune_dyn_beam(lat)  # defaults: limits INT_MAX -> untouched
    assert len(r["arcs"]) == lat.narcs and r["nstates"] == lat.nstates and r["iters"] == 0
    r2 = ora.prune_dyn_beam(lat, max_arcs=lat.narcs // 4, max_states=lat.nstates)
    assert 0 < len(r2["arcs"]) <= lat.narcs // 4 or r2["beam"] <= 1e-3
    assert r2["beam"] < r2["beam0"]
    # survivors keep their relative order
    idx = [a[0] for a in r2["arcs"]]
    assert idx == sorted(idx)


# ---- independent pins for the V rows (no golden in the reference) ------------------------
@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("flags", [dict(), dict(acoustic_scale=0.3, insertion_penalty=0.5), dict(graph_scale=2.0)])
def test_oracle_best_path2_vs_all_paths(klu, ora, seed, flags):
    """latbin/lattice-best-path2.cc:122-199 by its definition: pad all label sequences to the
    longest one, posterior of (label, position) from the list of all paths, expected position
    loss of every label sequence, arg-min.  The restatement (length unfolding + padding chain +
    forward-backward + float shortest path) must pick the same sequence whenever the runner-up
    is further away than float noise, and report the same cost."""
    checked = 0
    for lat in klu.synth_batch("tiny", 5, seed=700 + seed).lattices():
        labels, cost, margin = ora.brute_best_path2(lat, **flags)
        got_labels, got_cost = ora.best_path2(lat, **flags)
        assert abs(got_cost - cost) <= 1e-5 * max(1.0, cost)
        if margin > 1e-4:
            assert got_labels == labels
            checked += 1
    assert checked >= 3


def test_oracle_best_path2_vs_all_paths_padded_sequence(klu, ora):
    # two label sequences of different length on a diamond: the shorter one is padded with kNoLabel
    arcs = [(0, 1, 5, 1.0, 0.0, 1), (0, 2, 6, 0.25, 0.0, 1), (1, 3, 7, 0.5, 0.0, 1), (2, 3, 0, 0.5, 0.0, 1),
            (3, 4, 8, 0.0, 0.0, 1)]
    lat = klu.make_lattice("diamond", 5, arcs, {4: (0.0, 0.0)})
    labels, cost, margin = ora.brute_best_path2(lat)
    got_labels, got_cost = ora.best_path2(lat)
    assert abs(got_cost - cost) < 1e-6
    assert margin > 1e-3 and got_labels == labels == [6, 8]


PRUNE_FLAGS = [dict(max_arcs=40, max_states=30), dict(max_arcs=15), dict(max_states=12, beam_ratio=0.5),
               dict(acoustic_scale=0.1, graph_scale=0.7, insertion_penalty=0.5, max_arcs=60),
               dict(max_arcs=1, min_beam=0.5), dict()]


@pytest.mark.parametrize("seed", range(5))
@pytest.mark.parametrize("flags", PRUNE_FLAGS)
def test_oracle_prune_dyn_beam_vs_all_paths(klu, ora, seed, flags):
    """latbin/lattice-prune-dyn-beam.cc:27-90,166-184 by its definition: an arc survives beam b
    iff it lies on a path of cost <= best + b, the lattice's own beam is the largest such
    distance, the beam shrinks by --beam-ratio until the limits hold.  Surviving arcs, their new
    state ids and their output weights must be identical; the beams agree to rounding (the
    sweeps add the same costs in another order)."""
    import math
    checked = 0
    for lat in klu.synth_batch("tiny", 4, seed=800 + seed).lattices():
        want = ora.brute_prune_dyn_beam(lat, **flags)
        got = ora.prune_dyn_beam(lat, **flags)
        assert got["iters"] == want["iters"]
        assert math.isclose(got["beam0"], want["beam0"], rel_tol=1e-12, abs_tol=1e-12)
        assert math.isclose(got["beam"], want["beam"], rel_tol=1e-12, abs_tol=1e-12)
        if want["margin"] > 1e-9:
            assert got["nstates"] == want["nstates"]
            assert got["arcs"] == want["arcs"] and got["finals"] == want["finals"]
            checked += 1
    assert checked >= 3


@pytest.mark.parametrize("seed", range(8))
def test_oracle_top_order_properties(klu, ora, seed):
    """[ext] fst::TopSort restated as a recursive depth-first visit: the order is a
    topological one (a permutation with every arc going up)."""
    from test_topsort import _random_dag
    rng = np.random.RandomState(seed)
    lat = _random_dag(klu, rng, 4 + 5 * seed)
    order = ora.top_order(lat)
    assert sorted(order) == list(range(lat.nstates))
    assert all(order[s] < order[d] for s, d in zip(lat.src.tolist(), lat.dst.tolist()))


@pytest.mark.parametrize("seed", range(5))
@pytest.mark.parametrize("flags", [dict(beam=0.05), dict(beam=0.5), dict(beam=3.0), dict(),
                                   dict(beam=0.2, acoustic_scale=0.1, graph_scale=0.7, insertion_penalty=0.5)])
def test_oracle_prune_arcs_vs_all_paths(klu, ora, seed, flags):
    """latbin/lattice-prune-arcs.cc:34-84 by its definition (an arc's cost-through is -log of the
    mass of the paths through it; ascending sort; accumulate; put back the arcs from the cut on;
    keep what still lies on a path of kept arcs) against the restatement through alpha/beta."""
    checked = 0
    for lat in klu.synth_batch("tiny", 4, seed=900 + seed).lattices():
        want = ora.brute_prune_arcs(lat, **flags)
        got = ora.prune_arcs(lat, **flags)
        assert abs(got["cutoff"] - want["cutoff"]) <= 1e-9 * max(1.0, abs(want["cutoff"])) or want["cutoff"] == got["cutoff"]
        if want["margin"] > 1e-9:
            assert got["first_kept"] == want["first_kept"] and got["nstates"] == want["nstates"]
            assert got["arcs"] == want["arcs"] and got["finals"] == want["finals"]
            checked += 1
    assert checked >= 3


def test_oracle_prune_arcs_default_beam_keeps_everything(klu, ora):
    lat = klu.synth_batch("small", 1, seed=3)[0]
    r = ora.prune_arcs(lat)
    assert r["first_kept"] == 0 and r["nstates"] == lat.nstates and len(r["arcs"]) == lat.narcs
    assert sorted(a[0] for a in r["arcs"]) == list(range(lat.narcs))
    # a state's arcs come back in ascending cost-through order (AddArc appends), not in stored order
    assert [a[0] for a in r["arcs"]] != list(range(lat.narcs))
