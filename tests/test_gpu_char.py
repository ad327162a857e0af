"""lattice-char-index-position on the GPU against the CPU oracle and the README golden
(kwsbin2/README.md:232, pinned for the oracle in tests/test_oracle_golden.py)."""
import os
import subprocess

import numpy as np
import pytest

from util import GOLD, TOL, assert_rows_match, char_lattice, goldens

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check(got, want, what):
    # rows: (string, pos, t0, t1, logp); key = (string, pos), (t0, t1) exact, logp within TOL
    assert_rows_match(got, want, 2, what=what)


def test_readme_char_lattice(klu, ora, engine):
    lat = klu.read_text_ark(os.path.join(GOLD, "lattice.char.ark.txt"))[0]
    engine.load(klu.LatticeBatch.from_lattices([lat]))
    got = engine.char_position([28])[0]
    want = ora.char_position(lat, [28])
    _check(got, want, "README char lattice")
    assert [r[:4] for r in got] == [r[:4] for r in want]  # same order too (no near-ties in this example)


def test_readme_char_cli_matches_golden_within_reference_noise():
    # the README values carry the reference's float32 / determinize-delta noise (<= 5.9e-5)
    tool = os.path.join(ROOT, "kaldi-lattice-utils_b200", "bin", "lattice-char-index-position")
    r = subprocess.run([tool, "28", "ark:" + os.path.join(GOLD, "lattice.char.ark.txt"), "ark,t:-"],
                       capture_output=True)
    assert r.returncode == 0, r.stderr.decode()

    def parse(line):
        key, rest = line.strip().split(" ", 1)
        rows = []
        for ent in rest.split(";"):
            f = ent.split()
            if f:
                rows.append((f[0], int(f[1]), int(f[2]), int(f[3]), float(f[4])))
        return key, rows
    gk, got = parse(r.stdout.decode())
    wk, want = parse(goldens()["char_position"])
    assert gk == wk
    assert [x[:4] for x in got] == [x[:4] for x in want]
    assert max(abs(a[4] - b[4]) for a, b in zip(got, want)) <= TOL


CASES = [dict(), dict(acoustic_scale=0.5), dict(graph_scale=0.8, insertion_penalty=0.3), dict(beam=4.0),
         dict(nbest=7), dict(nbest=100000)]


@pytest.mark.parametrize("flags", CASES)
def test_char_position_parity_random(klu, ora, engine, flags):
    rng = np.random.RandomState(1234)
    lats = [char_lattice(klu, rng, "c%d" % i, nwords=int(rng.randint(1, 5))) for i in range(12)]
    lats.insert(4, klu.make_lattice("empty", 0, [], {}))
    engine.load(klu.LatticeBatch.from_lattices(lats))
    got = engine.char_position([1], **flags)
    for l, lat in enumerate(lats):
        want = ora.char_position(lat, [1], **flags)
        if flags.get("nbest", 100) < 50:
            # the cut falls between near-equal scores only by accident: compare the kept prefix as a set
            assert len(got[l]) == len(want)
            _check(got[l], want, "char lat %d %r" % (l, flags))
        else:
            _check(got[l], want, "char lat %d %r" % (l, flags))


def test_char_position_other_groups_and_epsilons(klu, ora, engine):
    rng = np.random.RandomState(99)
    lats = [char_lattice(klu, rng, "p%d" % i, nwords=3, punct=(40, 41), eps_prob=0.08) for i in range(10)]
    engine.load(klu.LatticeBatch.from_lattices(lats))
    got = engine.char_position([1], other_groups=[[40], [41]], nbest=100000)
    for l, lat in enumerate(lats):
        _check(got[l], ora.char_position(lat, [1], other_groups=[[40], [41]], nbest=100000), "punct lat %d" % l)
    got = engine.char_position([1], other_groups=[[40, 41]], nbest=100000)
    for l, lat in enumerate(lats):
        _check(got[l], ora.char_position(lat, [1], other_groups=[[40, 41]], nbest=100000), "punct2 lat %d" % l)


# ---- lattice-char-index-segment ----------------------------------
def _check_seg(got, want, what):
    # rows (string, t0, t1, logp).  Sub-paths with the same string and (t0, t1) but
    # different intermediate frame tags are separate rows (SURVEY.md 8c hazard 9), so
    # rows are compared as sorted multisets: keys exact, values within TOL
    assert len(got) == len(want), what
    for g, w in zip(sorted(got, key=lambda r: (r[:3], -r[3])), sorted(want, key=lambda r: (r[:3], -r[3]))):
        assert g[:3] == w[:3], what
        assert abs(g[3] - w[3]) <= TOL, what
    for a, b in zip(got[:-1], got[1:]):
        assert a[3] >= b[3]


def test_readme_char_segment(klu, ora, engine):
    lat = klu.read_text_ark(os.path.join(GOLD, "lattice.char.ark.txt"))[0]
    engine.load(klu.LatticeBatch.from_lattices([lat]))
    got = engine.char_segment([28])[0]
    want = ora.char_segment(lat, [28])
    _check_seg(got, want, "README char lattice")
    assert [r[:3] for r in got] == [r[:3] for r in want]


def test_readme_char_segment_cli():
    tool = os.path.join(ROOT, "kaldi-lattice-utils_b200", "bin", "lattice-char-index-segment")
    r = subprocess.run([tool, "28", "ark:" + os.path.join(GOLD, "lattice.char.ark.txt"), "ark,t:-"],
                       capture_output=True)
    assert r.returncode == 0, r.stderr.decode()

    def parse(line):
        key, rest = line.strip().split(" ", 1)
        rows = []
        for ent in rest.split(";"):
            f = ent.split()
            if f:
                rows.append((f[0], int(f[1]), int(f[2]), float(f[3])))
        return key, rows
    gk, got = parse(r.stdout.decode())
    wk, want = parse(goldens()["char_segment"])
    assert gk == wk and [x[:3] for x in got] == [x[:3] for x in want]
    assert max(abs(a[3] - b[3]) for a, b in zip(got, want)) <= TOL


@pytest.mark.parametrize("flags", [dict(nbest=100000), dict(acoustic_scale=0.5, nbest=100000), dict(beam=4.0, nbest=100000)])
def test_char_segment_parity_random(klu, ora, engine, flags):
    rng = np.random.RandomState(4321)
    lats = [char_lattice(klu, rng, "s%d" % i, nwords=int(rng.randint(1, 5)), punct=(40,), eps_prob=0.05)
            for i in range(12)]
    engine.load(klu.LatticeBatch.from_lattices(lats))
    got = engine.char_segment([1], other_groups=[[40]], **flags)
    for l, lat in enumerate(lats):
        _check_seg(got[l], ora.char_segment(lat, [1], other_groups=[[40]], **flags), "seg lat %d %r" % (l, flags))


def test_char_segment_c5(klu, ora, engine):
    batch = klu.synth_batch("c5", 2, seed=9)
    engine.load(batch)
    got = engine.char_segment([1], nbest=100000)
    for l, lat in enumerate(batch.lattices()):
        _check_seg(got[l], ora.char_segment(lat, [1], nbest=100000), "c5 seg lat %d" % l)
