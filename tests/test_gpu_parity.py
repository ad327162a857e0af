"""Parity of the CUDA path (through the C ABI) against the CPU oracle."""
import os

import numpy as np
import pytest

from util import GOLD, TOL, assert_rows_match, goldens

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def word_lat(klu):
    return klu.read_text_ark(os.path.join(GOLD, "lattice.ark.txt"))[0]


def _load(klu, engine, lats):
    engine.load(klu.LatticeBatch.from_lattices(lats))


def test_readme_segment_text(klu, engine, word_lat):
    _load(klu, engine, [word_lat])
    assert klu.format_tuples("lat1", engine.segment()[0]).strip() == goldens()["segment"]


def test_readme_position_text(klu, engine, word_lat):
    _load(klu, engine, [word_lat])
    assert klu.format_tuples("lat1", engine.position()[0]).strip() == goldens()["position"]


def test_readme_utterance_text(klu, engine, word_lat):
    _load(klu, engine, [word_lat])
    assert klu.format_tuples("lat1", engine.utterance()[0]).strip() == goldens()["utterance"]


def test_fwd_bwd_matches_oracle(klu, ora, engine):
    batch = klu.synth_batch("small", 8, seed=11)
    engine.load(batch)
    engine.run(klu.FWD_BWD)
    al, be, tot = engine.fetch_fwd_bwd()
    for l, lat in enumerate(batch.lattices()):
        seg = ora.run(ora.SEGMENT, lat)
        assert abs(tot[l] - seg.ds0) < 1e-9


SHAPES = [("tiny", 24, 5), ("small", 12, 6)]
FLAGS = [dict(), dict(acoustic_scale=0.1), dict(graph_scale=0.7, insertion_penalty=0.5),
         dict(acoustic_scale=0.3, beam=6.0)]


@pytest.mark.parametrize("shape,n,seed", SHAPES)
@pytest.mark.parametrize("flags", FLAGS)
def test_segment_parity(klu, ora, engine, shape, n, seed, flags):
    batch = klu.synth_batch(shape, n, seed=seed)
    engine.load(batch)
    got = engine.segment(**flags)
    for l, lat in enumerate(batch.lattices()):
        assert_rows_match(got[l], ora.segment(lat, **flags), 3, what="segment lat %d" % l)


@pytest.mark.parametrize("shape,n,seed", SHAPES)
@pytest.mark.parametrize("flags", FLAGS)
def test_position_parity(klu, ora, engine, shape, n, seed, flags):
    batch = klu.synth_batch(shape, n, seed=seed + 50)
    engine.load(batch)
    got = engine.position(**flags)
    for l, lat in enumerate(batch.lattices()):
        assert_rows_match(got[l], ora.position(lat, **flags), 2, what="position lat %d" % l)


@pytest.mark.parametrize("shape,n,seed", SHAPES)
@pytest.mark.parametrize("flags", FLAGS[:3])
def test_frame_post_parity(klu, ora, engine, shape, n, seed, flags):
    batch = klu.synth_batch(shape, n, seed=seed + 90)
    engine.load(batch)
    got = engine.frame_post(**flags)
    for l, lat in enumerate(batch.lattices()):
        want = ora.frame_post(lat, **flags)
        assert len(got[l]) == len(want)
        for k, (g, w) in enumerate(zip(got[l], want)):
            assert_rows_match([(a, b) for a, b in g], [(a, b) for a, b in w], 1, what="frame %d lat %d" % (k, l))


@pytest.mark.parametrize("shape,n,seed", SHAPES)
@pytest.mark.parametrize("flags", FLAGS)
def test_utterance_parity(klu, ora, engine, shape, n, seed, flags):
    batch = klu.synth_batch(shape, n, seed=seed + 130)
    engine.load(batch)
    got = engine.utterance(**flags)
    for l, lat in enumerate(batch.lattices()):
        assert_rows_match(got[l], ora.utterance(lat, **flags), 1, what="utterance lat %d" % l)


def test_chunked_execution_matches_single_chunk(klu, ora, engine, monkeypatch):
    # tiny scratch budget -> many chunks; results must not change
    batch = klu.synth_batch("tiny", 16, seed=321)
    engine.load(batch)
    ref = dict(seg=engine.segment(), pos=engine.position(), fp=engine.frame_post(), utt=engine.utterance())
    monkeypatch.setenv("KLU_ENTRY_BUDGET", "150")
    assert engine.segment() == ref["seg"]
    assert engine.position() == ref["pos"]
    assert engine.frame_post() == ref["fp"]
    assert engine.utterance() == ref["utt"]


def test_include_exclude_words(klu, ora, engine):
    batch = klu.synth_batch("tiny", 6, seed=77)
    engine.load(batch)
    for flags in (dict(include_words=[1, 2, 3]), dict(exclude_words=[1, 4]), dict(include_words=[2], exclude_words=[2])):
        got = engine.segment(**flags)
        gotp = engine.position(**flags)
        gotu = engine.utterance(**flags)
        for l, lat in enumerate(batch.lattices()):
            assert_rows_match(got[l], ora.segment(lat, **flags), 3)
            assert_rows_match(gotp[l], ora.position(lat, **flags), 2)
            assert_rows_match(gotu[l], ora.utterance(lat, **flags), 1)


def test_empty_and_ragged_batch(klu, ora, engine):
    lats = klu.synth_batch("tiny", 3, seed=5).lattices()
    empty = klu.make_lattice("empty", 0, [], {})
    single = klu.make_lattice("single", 1, [], {0: (0.5, 0.25)})
    engine.load(klu.LatticeBatch.from_lattices([lats[0], empty, lats[1], single, lats[2]]))
    got = engine.segment()
    assert got[1] == [] and got[3] == []
    for i, j in ((0, 0), (2, 1), (4, 2)):
        assert_rows_match(got[i], ora.segment(lats[j]), 3)
    fp = engine.frame_post()
    assert fp[1] == [] and fp[3] == []


def test_rejects_unsorted_lattice(klu, engine):
    bad = klu.make_lattice("bad", 3, [(0, 2, 1, 0.1, 0.2, 1), (2, 1, 2, 0.1, 0.2, 1)], {1: (0.0, 0.0)})
    with pytest.raises(klu.KluError):
        engine.load(klu.LatticeBatch.from_lattices([bad]))


# ---- lattice-best-path2 ---------------------------------------------------------
def test_best_path2_readme_lattice(klu, engine, word_lat):
    _load(klu, engine, [word_lat])
    (labels, cost), = engine.best_path2()
    assert labels == [2, 3, 5, 2, 6, 7, 8]
    assert abs(cost - 0.4) < 1e-6


@pytest.mark.parametrize("shape,n,seed", SHAPES)
@pytest.mark.parametrize("flags", FLAGS[:3])
def test_best_path2_parity(klu, ora, engine, shape, n, seed, flags):
    batch = klu.synth_batch(shape, n, seed=seed + 170)
    engine.load(batch)
    got = engine.best_path2(**flags)
    for l, lat in enumerate(batch.lattices()):
        labels, cost = ora.best_path2(lat, **flags)
        assert got[l][0] == labels, "best-path2 labels, lattice %d" % l      # bit-exact label sequence
        assert abs(got[l][1] - cost) <= 1e-4 * max(1.0, abs(cost))


# ---- lattice-to-word-position-post ---------------------------
@pytest.mark.parametrize("shape,n,seed", SHAPES)
@pytest.mark.parametrize("flags", [dict(), dict(acoustic_scale=0.1), dict(graph_scale=0.7, insertion_penalty=0.5)])
def test_position_post_parity(klu, ora, engine, shape, n, seed, flags):
    batch = klu.synth_batch(shape, n, seed=seed + 77)
    engine.load(batch)
    got = engine.position_post(**flags)
    for l, lat in enumerate(batch.lattices()):
        want = ora.position_post(lat, **flags)
        assert len(got[l]) == len(want), "number of positions, lattice %d" % l
        for k, (g, w) in enumerate(zip(got[l], want)):
            assert_rows_match(g, w, 1, what="position-post lat %d pos %d" % (l, k))


def test_position_post_readme_lattice(klu, ora, engine, word_lat):
    _load(klu, engine, [word_lat])
    got = engine.position_post()[0]
    want = ora.position_post(word_lat)
    assert [[w for w, _ in pos] for pos in got] == [[w for w, _ in pos] for pos in want]
    for g, w in zip(got, want):
        for (_, a), (_, b) in zip(g, w):
            assert a == b  # float32 values, bit-identical on this chain/diamond lattice


# ---- lattice-to-transcript-length-dist -----------------------
@pytest.mark.parametrize("shape,n,seed", SHAPES)
@pytest.mark.parametrize("flags", [dict(), dict(acoustic_scale=0.1, insertion_penalty=0.5)])
def test_length_dist_parity(klu, ora, engine, shape, n, seed, flags):
    batch = klu.synth_batch(shape, n, seed=seed + 5)
    lats = batch.lattices() + [klu.make_lattice("empty", 0, [], {})]
    engine.load(klu.LatticeBatch.from_lattices(lats))
    got = engine.length_dist(**flags)
    for l, lat in enumerate(lats):
        assert_rows_match(got[l], ora.length_dist(lat, **flags), 1, what="length-dist lat %d" % l)


def test_length_dist_readme_lattice(klu, engine, word_lat):
    _load(klu, engine, [word_lat])
    assert engine.length_dist()[0] == [(7, 0.0)]


# ---- lattice-prune-dyn-beam -----------------------------------------------------
def _check_prune(got, want, lat):
    assert got["nstates"] == want["nstates"]
    assert [a[:4] for a in got["arcs"]] == [a[:4] for a in want["arcs"]]          # surviving arc set, bit-exact
    assert [a[4:] for a in got["arcs"]] == [a[4:] for a in want["arcs"]]          # output weights, bit-exact floats
    assert got["finals"] == want["finals"]
    assert got["beam0"] == want["beam0"] and got["beam"] == want["beam"]           # doubles, bit-exact


@pytest.mark.parametrize("shape,n,seed", SHAPES)
@pytest.mark.parametrize("flags", [dict(), dict(max_arcs=40, max_states=30), dict(max_arcs=15),
                                   dict(max_states=12, beam_ratio=0.5),
                                   dict(acoustic_scale=0.1, graph_scale=0.7, insertion_penalty=0.5, max_arcs=60),
                                   dict(max_arcs=1, min_beam=0.5)])
def test_prune_dyn_beam_parity(klu, ora, engine, shape, n, seed, flags):
    batch = klu.synth_batch(shape, n, seed=seed + 210)
    engine.load(batch)
    got = engine.prune_dyn_beam(**flags)
    for l, lat in enumerate(batch.lattices()):
        _check_prune(got[l], ora.prune_dyn_beam(lat, **flags), lat)


def test_prune_then_best_path_pipeline(klu, ora, engine):
    # configs[2]: prune-dyn-beam piped into best-path2
    batch = klu.synth_batch("small", 6, seed=999)
    engine.load(batch)
    flags = dict(max_arcs=200, max_states=120)
    pruned = engine.prune_dyn_beam(**flags)
    lats2 = []
    for l, lat in enumerate(batch.lattices()):
        p = pruned[l]
        arcs = [(s, d, lab, g, a, int(lat.dur[i])) for i, s, d, lab, g, a in p["arcs"]]
        finals = {s: (g, a) for s, g, a in p["finals"]}
        lats2.append(klu.make_lattice(lat.key, p["nstates"], arcs, finals))
    engine.load(klu.LatticeBatch.from_lattices(lats2))
    got = engine.best_path2()
    for l, lat in enumerate(lats2):
        labels, cost = ora.best_path2(lat)
        assert got[l][0] == labels


# ---- frame-synchronous kernel vs the generic emit/sort/reduce pipeline ------------
def test_frame_post_fast_path_equals_generic(klu, engine, monkeypatch):
    batch = klu.synth_batch("small", 10, seed=4242)
    engine.load(batch)
    fast = engine.frame_post(acoustic_scale=0.1)
    monkeypatch.setenv("KLU_GENERIC_FRAME_POST", "1")
    generic = engine.frame_post(acoustic_scale=0.1)
    assert len(fast) == len(generic)
    for a, b in zip(fast, generic):
        assert len(a) == len(b)
        for fa, fb in zip(a, b):
            assert_rows_match(fa, fb, 1, tol=1e-6)


def test_frame_post_many_words_per_frame(klu, ora, engine):
    # > 256 distinct words alive in one frame: the per-warp hash table overflows and
    # the kernel falls back to word-hash partitions + a global-memory sort
    rng = np.random.RandomState(0)
    n = 900
    arcs = [(0, 1, 1000 + i, float(rng.uniform(0, 5)), float(rng.uniform(0, 5)), 2) for i in range(n)]
    arcs += [(0, 1, 1000 + i, float(rng.uniform(0, 5)), 0.25, 2) for i in range(0, n, 3)]   # duplicates to merge
    arcs += [(1, 2, 7, 0.5, 0.5, 1)]
    lat = klu.make_lattice("wide", 3, arcs, {2: (0.0, 0.0)})
    engine.load(klu.LatticeBatch.from_lattices([lat]))
    got = engine.frame_post()[0]
    want = ora.frame_post(lat)
    assert len(got) == len(want) == 3
    for g, w in zip(got, want):
        assert_rows_match(g, w, 1)


# ---- input without the per-arc source array ---------------------------------------
def test_load_from_state_arc_counts(klu, engine, monkeypatch):
    batch = klu.synth_batch("small", 7, seed=31)
    engine.load(batch)
    want = dict(seg=engine.segment(acoustic_scale=0.3), fp=engine.frame_post(), pr=engine.prune_dyn_beam(max_arcs=150))
    counts = batch.state_num_arcs()
    for host in (False, True):
        if host:
            monkeypatch.setenv("KLU_HOST_PACKER", "1")
        engine.load(batch, state_num_arcs=counts)
        got = dict(seg=engine.segment(acoustic_scale=0.3), fp=engine.frame_post(), pr=engine.prune_dyn_beam(max_arcs=150))
        assert got == want
    bad = counts.copy()
    bad[0] += 1
    with pytest.raises(klu.KluError):
        engine.load(batch, state_num_arcs=bad)


def test_load_from_compact_arc_arrays(klu, engine, monkeypatch):
    """klu_lattices.arc_dur_u8 / arc_dst_delta_u16 (15 instead of 20 bytes per arc on the way up)."""
    batch = klu.synth_batch("small", 7, seed=32)
    engine.load(batch)
    want = dict(seg=engine.segment(acoustic_scale=0.3), pos=engine.position(), fp=engine.frame_post())
    dur8, d16 = klu.binding.compact_arcs(batch)
    assert dur8 is not None and d16 is not None
    for host in (False, True):
        if host:
            monkeypatch.setenv("KLU_HOST_PACKER", "1")
        for kw in (dict(dur_u8=dur8), dict(dst_delta_u16=d16), dict(dur_u8=dur8, dst_delta_u16=d16),
                   dict(dur_u8=dur8, dst_delta_u16=d16, state_num_arcs=batch.state_num_arcs())):
            engine.load(batch, **kw)
            got = dict(seg=engine.segment(acoustic_scale=0.3), pos=engine.position(), fp=engine.frame_post())
            assert got == want
    bad = d16.copy()
    bad[3] = 0  # dst == src: not topologically sorted
    with pytest.raises(klu.KluError):
        engine.load(batch, dst_delta_u16=bad)


def test_load_from_16_bit_labels(klu, engine, monkeypatch):
    """klu_lattices.arc_label_u16 (13 bytes per arc on the way up with the other compact forms)."""
    batch = klu.synth_batch("small", 7, seed=33)
    engine.load(batch)
    want = dict(seg=engine.segment(acoustic_scale=0.3), pos=engine.position(), fp=engine.frame_post(),
                utt=engine.utterance(), bp=engine.best_path2())
    lab16 = klu.binding.compact_labels(batch)
    dur8, d16 = klu.binding.compact_arcs(batch)
    assert lab16 is not None
    for host in (False, True):
        if host:
            monkeypatch.setenv("KLU_HOST_PACKER", "1")
        for kw in (dict(label_u16=lab16), dict(label_u16=lab16, dur_u8=dur8, dst_delta_u16=d16,
                                              state_num_arcs=batch.state_num_arcs())):
            engine.load(batch, **kw)
            got = dict(seg=engine.segment(acoustic_scale=0.3), pos=engine.position(), fp=engine.frame_post(),
                       utt=engine.utterance(), bp=engine.best_path2())
            assert got == want
    wide = klu.make_lattice("wide-label", 2, [(0, 1, 70000, 0.5, 0.5, 1)], {1: (0.0, 0.0)})
    assert klu.binding.compact_labels(klu.LatticeBatch.from_lattices([wide])) is None


@pytest.mark.parametrize("generic", [False, True])
def test_frame_post_rows_by_frame_offsets(klu, engine, monkeypatch, generic):
    """klu_fetch_frame_post_csr: the same rows as klu_fetch_frame_post, per-frame offsets instead of a frame
    column (empty lattices and frames without a word included)."""
    lats = klu.synth_batch("small", 6, seed=34).lattices() + klu.synth_batch("tiny", 9, seed=35).lattices()
    lats.insert(2, klu.make_lattice("empty", 0, [], {}))
    lats.insert(5, klu.make_lattice("eps-only", 3, [(0, 1, 0, 0.5, 0.5, 2), (1, 2, 0, 0.1, 0.2, 1)], {2: (0.0, 0.0)}))
    if generic:
        monkeypatch.setenv("KLU_GENERIC_FRAME_POST", "1")
    engine.load(klu.LatticeBatch.from_lattices(lats))
    for flags in (dict(), dict(acoustic_scale=0.1)):
        assert engine.frame_post_csr(**flags) == engine.frame_post(**flags)


# ---- device packer vs host packer ----------------------------------------------
def test_gpu_packer_equals_host_packer(klu, engine, monkeypatch):
    lats = klu.synth_batch("small", 9, seed=2024).lattices()
    lats.insert(3, klu.make_lattice("empty", 0, [], {}))
    lats.insert(5, klu.make_lattice("single", 1, [], {0: (0.5, 0.25)}))
    batch = klu.LatticeBatch.from_lattices(lats)

    def run_all():
        engine.load(batch)
        engine.run(klu.FWD_BWD)
        return dict(fb=[x.tolist() for x in engine.fetch_fwd_bwd()], seg=engine.segment(acoustic_scale=0.3),
                    pos=engine.position(), utt=engine.utterance(), fp=engine.frame_post(),
                    bp=engine.best_path2(), pr=engine.prune_dyn_beam(max_arcs=150))

    gpu = run_all()
    monkeypatch.setenv("KLU_HOST_PACKER", "1")
    host = run_all()
    assert gpu["seg"] == host["seg"] and gpu["pos"] == host["pos"] and gpu["utt"] == host["utt"]
    assert gpu["fp"] == host["fp"] and gpu["bp"] == host["bp"] and gpu["pr"] == host["pr"]
    assert gpu["fb"] == host["fb"]


# ---- edge cases of the frame-synchronous path --------------------------------------
def test_frame_post_underflowing_group_is_redone_exactly(klu, ora, engine):
    # word 9 only occurs on an arc 2000 nats worse than the best path: exp(-2000) underflows
    # the f64 posterior, so its group goes through the exact log-domain path
    arcs = [(0, 1, 5, 1.0, 0.5, 2), (0, 1, 9, 2000.0, 0.5, 2), (0, 1, 6, 1.5, 0.25, 2), (1, 2, 7, 0.5, 0.5, 3),
            (1, 2, 9, 2500.0, 0.0, 3)]
    lat = klu.make_lattice("under", 3, arcs, {2: (0.0, 0.0)})
    engine.load(klu.LatticeBatch.from_lattices([lat]))
    got = engine.frame_post()[0]
    want = ora.frame_post(lat)
    assert len(got) == len(want) == 5
    for g, w in zip(got, want):
        assert [x[0] for x in g] == [x[0] for x in w]
        for (_, a), (_, b) in zip(g, w):
            assert abs(a - b) <= 1e-4 * max(1.0, abs(b))
    assert got[0][-1][0] == 9 and -2001.0 < got[0][-1][1] < -1997.0


def test_frame_post_ties_and_near_ties(klu, ora, engine):
    # exact ties (identical weights, different words) order by word; values that differ only in
    # the low mantissa bits of the float log-posterior exercise the 64-bit re-sort
    rng = np.random.RandomState(3)
    arcs = []
    for w in range(40):
        arcs.append((0, 1, 100 + w, 1.0, 0.5, 1))                       # 40-way exact tie
    for w in range(40):
        arcs.append((0, 1, 200 + w, 3.0 + 1e-7 * rng.randint(0, 50), 0.5, 1))  # near-ties
    arcs.append((1, 2, 7, 0.5, 0.5, 1))
    lat = klu.make_lattice("ties", 3, arcs, {2: (0.0, 0.0)})
    engine.load(klu.LatticeBatch.from_lattices([lat]))
    got = engine.frame_post()[0]
    want = ora.frame_post(lat)
    assert [x[0] for x in got[0][:40]] == list(range(100, 140))
    for g, w in zip(got, want):
        assert_rows_match(g, w, 1, what="ties")
        for a, b in zip(g[:-1], g[1:]):  # strictly the reference's comparator on OUR floats
            assert a[1] > b[1] or (a[1] == b[1] and a[0] < b[0])


def test_index_order_with_ties_and_near_ties(klu, ora, engine):
    """The index tools sort by (double logp desc, key asc).  The order sort looks at the
    high half of the key only and settles what agrees there afterwards: exact ties
    must come out in key order, near-ties (1e-9 .. 1e-13 apart) in value order."""
    arcs = []
    for w in range(30):
        arcs.append((0, 1, 100 + w, 1.0, 0.5, 1))                          # 30-way exact tie
    for w in range(30):
        arcs.append((0, 1, 200 + w, 2.0, 0.5 + 2.0 ** -20 * (29 - w), 1))  # float weights one ulp-ish apart
    for w in range(30):
        arcs.append((0, 2, 300 + w, 1.5, 0.25, 2))
    arcs.append((1, 3, 7, 0.5, 0.5, 2))
    arcs.append((2, 3, 8, 0.5, 0.5, 1))
    lat = klu.make_lattice("order", 4, arcs, {3: (0.0, 0.0)})
    engine.load(klu.LatticeBatch.from_lattices([lat]))
    for name, ncol in (("segment", 3), ("position", 2), ("utterance", 1)):
        got, want = getattr(engine, name)()[0], getattr(ora, name)(lat)
        assert [r[:ncol] for r in got] == [r[:ncol] for r in want], name
        for a, b in zip(got[:-1], got[1:]):  # the comparator on OUR doubles
            assert a[-1] > b[-1] or (a[-1] == b[-1] and a[:ncol] < b[:ncol]), (name, a, b)


def test_unreachable_and_dead_end_states(klu, ora, engine):
    # state 2 cannot be reached from the start (and has no arcs: CompactLatticeStateTimes
    # asserts otherwise), state 3 cannot reach a final state
    arcs = [(0, 1, 5, 1.0, 0.5, 1), (0, 3, 8, 0.5, 0.5, 1), (1, 4, 6, 0.5, 0.5, 1)]
    lat = klu.make_lattice("holes", 5, arcs, {4: (0.0, 0.0)})
    engine.load(klu.LatticeBatch.from_lattices([lat]))
    engine.run(klu.FWD_BWD)
    al, be, tot = engine.fetch_fwd_bwd()
    assert al[2] == -np.inf and be[3] == -np.inf and np.isfinite(al[[0, 1, 3, 4]]).all()
    assert_rows_match(engine.segment()[0], ora.segment(lat), 3, what="holes segment")
    got, want = engine.frame_post()[0], ora.frame_post(lat)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert_rows_match(g, w, 1, what="holes frame-post")


# ---- (word, position) cells vs the generic one-entry-per-(arc, length) pipeline -----------
@pytest.mark.parametrize("shape,n,seed", SHAPES)
def test_position_cells_equal_generic_pipeline(klu, engine, monkeypatch, shape, n, seed):
    batch = klu.synth_batch(shape, n, seed=seed + 400)
    engine.load(batch)
    flags = dict(acoustic_scale=0.3, beam=7.0)
    new = dict(pos=engine.position(**flags), pp=engine.position_post(acoustic_scale=0.3))
    monkeypatch.setenv("KLU_GENERIC_POSITION", "1")
    old = dict(pos=engine.position(**flags), pp=engine.position_post(acoustic_scale=0.3))
    for a, b in zip(new["pos"], old["pos"]):
        assert_rows_match(a, b, 2, tol=1e-9, what="cells vs generic")
    for a, b in zip(new["pp"], old["pp"]):
        assert len(a) == len(b)
        for fa, fb in zip(a, b):
            assert_rows_match(fa, fb, 1, tol=1e-6, what="cells vs generic, position-post")


# ---- segment index by start-frame buckets vs the generic sort-by-key pipeline ------------------
@pytest.mark.parametrize("shape,n,seed", SHAPES)
@pytest.mark.parametrize("flags", [dict(), dict(acoustic_scale=0.3, beam=6.0), dict(include_words=[3, 4, 7, 11, 40]),
                                   dict(exclude_words=[2, 5], graph_scale=0.5, insertion_penalty=0.3)])
def test_segment_buckets_equal_generic_pipeline(klu, engine, monkeypatch, shape, n, seed, flags):
    batch = klu.synth_batch(shape, n, seed=seed + 700)
    # a bucket over the cap (300 parallel arcs out of one state) sends the whole batch down the generic path
    wide = klu.make_lattice("wide", 3, [(0, 1, 1 + (i % 9), 0.1 * i, 0.05 * (i % 7), 2) for i in range(300)]
                            + [(1, 2, 4, 0.5, 0.5, 1)], {2: (0.0, 0.0)})
    # labels near 2^31 with a long arc: 32-bit sort words no longer fit, 64-bit ones do
    big = klu.make_lattice("big-labels", 4, [(0, 1, 2**31 - 2, 0.5, 0.25, 3), (0, 1, 2**31 - 2, 0.75, 0.5, 3),
                                             (0, 1, 2**30 + 5, 0.1, 0.2, 3), (1, 2, 7, 0.3, 0.3, 200), (1, 2, 7, 0.2, 0.1, 200),
                                             (2, 3, 2**31 - 2, 0.0, 0.1, 1)], {3: (0.0, 0.0)})
    for lats in (batch.lattices(), batch.lattices()[:3] + [wide], batch.lattices()[:2] + [big]):
        monkeypatch.delenv("KLU_GENERIC_SEGMENT", raising=False)
        engine.load(klu.LatticeBatch.from_lattices(lats))
        new = engine.segment(**flags)
        monkeypatch.setenv("KLU_GENERIC_SEGMENT", "1")
        old = engine.segment(**flags)
        assert len(new) == len(old) == len(lats)
        for a, b in zip(new, old):
            assert a == b  # same keys, same order, bit-identical log-posteriors


# ---- lattice-prune-arcs ------------------------------------------------
@pytest.mark.parametrize("shape,n,seed", SHAPES)
@pytest.mark.parametrize("flags", [dict(), dict(beam=0.05), dict(beam=0.5), dict(beam=3.0),
                                   dict(beam=0.2, acoustic_scale=0.1, graph_scale=0.7, insertion_penalty=0.5)])
def test_prune_arcs_parity(klu, ora, engine, shape, n, seed, flags):
    batch = klu.synth_batch(shape, n, seed=seed + 610)
    lats = batch.lattices() + [klu.make_lattice("empty", 0, [], {}), klu.make_lattice("single", 1, [], {0: (0.5, 0.25)})]
    engine.load(klu.LatticeBatch.from_lattices(lats))
    got = engine.prune_arcs(**flags)
    for l, lat in enumerate(lats):
        want = ora.prune_arcs(lat, **flags)
        assert got[l]["first_kept"] == want["first_kept"], "lattice %d" % l
        assert got[l]["nstates"] == want["nstates"]
        assert got[l]["arcs"] == want["arcs"]        # surviving arcs in the order AddArc left them, float weights bit-exact
        assert got[l]["finals"] == want["finals"]


def test_prune_arcs_ties_keep_arc_order(klu, ora, engine):
    # arcs of exactly equal cost-through: the sort is stable (documented tie rule)
    arcs = [(0, 1, 5 + k, 1.0, 0.5, 1) for k in range(6)] + [(1, 2, 7, 0.5, 0.5, 1), (1, 2, 8, 0.5, 0.5, 1)]
    lat = klu.make_lattice("ties", 3, arcs, {2: (0.0, 0.0)})
    engine.load(klu.LatticeBatch.from_lattices([lat]))
    for beam in (0.1, 1.0, 2.5):
        got, want = engine.prune_arcs(beam=beam)[0], ora.prune_arcs(lat, beam=beam)
        assert got["arcs"] == want["arcs"] and got["first_kept"] == want["first_kept"]
