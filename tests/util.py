"""Shared helpers of the parity tests."""
import json
import math
import os

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-4  # |delta log-posterior| bound stated by BASELINE.json's north_star


def goldens():
    return json.load(open(os.path.join(GOLD, "README_goldens.json")))


def close(a, b, tol=TOL):
    if math.isinf(a) or math.isinf(b):
        return a == b
    return abs(a - b) <= tol


def assert_rows_match(got, want, nkey, tol=TOL, what=""):
    """Rows are tuples whose first nkey fields are exact keys and whose last field
    is a log-probability.  Keys (and any exact fields between) must match
    bit-exactly as a keyed map; values within tol; and the ORDER must agree
    wherever the reference order is decided by more than tol (near-ties may
    legitimately permute, SURVEY.md 7 'ordering parity')."""
    gm = {r[:nkey]: r for r in got}
    wm = {r[:nkey]: r for r in want}
    assert len(gm) == len(got), what + ": duplicate keys in result"
    assert set(gm) == set(wm), what + ": key sets differ: missing %s extra %s" % (
        sorted(set(wm) - set(gm))[:5], sorted(set(gm) - set(wm))[:5])
    for k, w in wm.items():
        g = gm[k]
        assert close(g[-1], w[-1], tol), what + ": value of %s: %r vs %r" % (k, g[-1], w[-1])
        assert g[nkey:-1] == w[nkey:-1], what + ": exact fields of %s: %r vs %r" % (k, g, w)
    # order: non-increasing values, and positions agree up to tolerance
    for i in range(1, len(got)):
        assert got[i - 1][-1] >= got[i][-1] or close(got[i - 1][-1], got[i][-1], 0.0), what + ": not sorted"
    for i, (g, w) in enumerate(zip(got, want)):
        if g[:nkey] != w[:nkey]:
            assert close(g[-1], w[-1], tol), what + ": order differs beyond tolerance at %d: %r vs %r" % (i, g, w)


def assert_cols_match(gkeys, gexact, gval, wkeys, wexact, wval, tol=TOL, what="", ordered=True):
    """assert_rows_match over numpy columns (for results with millions of rows): keys and
    exact fields identical as a keyed map, values within tol, and -- when `ordered` -- values
    non-increasing with the reference's order wherever it is decided by more than tol."""
    import numpy as np
    gkeys, wkeys = [np.asarray(k) for k in gkeys], [np.asarray(k) for k in wkeys]
    gval, wval = np.asarray(gval, dtype=np.float64), np.asarray(wval, dtype=np.float64)
    assert len(gval) == len(wval), what + ": %d rows, reference has %d" % (len(gval), len(wval))
    go, wo = np.lexsort(gkeys[::-1]), np.lexsort(wkeys[::-1])
    for g, w in zip(gkeys, wkeys):
        assert np.array_equal(g[go], w[wo]), what + ": key sets differ"
    if len(gval) > 1:
        same = np.ones(len(gval) - 1, bool)
        for g in gkeys:
            same &= g[go][1:] == g[go][:-1]
        assert not same.any(), what + ": duplicate keys in result"
    for g, w in zip(gexact, wexact):
        assert np.array_equal(np.asarray(g)[go], np.asarray(w)[wo]), what + ": exact fields differ"
    gv, wv = gval[go], wval[wo]
    inf = np.isinf(gv) | np.isinf(wv)
    assert np.array_equal(gv[inf], wv[inf]), what + ": infinite values differ"
    if (~inf).any():
        d = np.abs(gv[~inf] - wv[~inf]).max()
        assert d <= tol, what + ": values differ by %g" % d
    if ordered and len(gval) > 1:
        assert (np.diff(gval) <= 0).all(), what + ": not sorted"
        moved = np.zeros(len(gval), bool)
        for g, w in zip(gkeys, wkeys):
            moved |= g != w
        if moved.any():
            ok = np.abs(gval[moved] - wval[moved]) <= tol
            assert ok.all(), what + ": order differs beyond tolerance"


def char_lattice(klu, rng, key, nwords=4, alphabet=5, punct=(), eps_prob=0.0):
    """Synthetic HTR-like character lattice: a time-layered DAG, one frame per arc,
    1-2 states per frame and 1-2 labels per state pair; label 1 is the whitespace
    delimiter.  Every few frames ALL arcs carry a delimiter (so same-group runs stay
    short enough for the oracle's explicit sub-path enumeration), elsewhere a delimiter
    is one of the alternatives now and then (so word counts differ between paths)."""
    layers = [[0]]
    nstates = 1
    arcs = []
    t = 0
    for w in range(nwords):
        wlen = int(rng.randint(1, 5))
        for k in range(wlen + 1):
            forced_space = k == wlen and w + 1 < nwords
            if k == wlen and w + 1 == nwords:
                break
            nxt = [nstates + i for i in range(int(rng.randint(1, 3)))]
            nstates += len(nxt)
            for u in layers[-1]:
                for v in nxt:
                    if len(nxt) > 1 and len(layers[-1]) > 1 and rng.rand() < 0.3:
                        continue
                    for _ in range(int(rng.randint(1, 3))):
                        if forced_space:
                            lab = 1
                        else:
                            x = rng.rand()
                            if x < 0.12:
                                lab = 1
                            elif x < 0.12 + eps_prob:
                                lab = 0
                            elif punct and x < 0.25 + eps_prob:
                                lab = int(punct[rng.randint(len(punct))])
                            else:
                                lab = 2 + int(rng.randint(alphabet))
                        arcs.append((u, v, lab, float(rng.uniform(0, 4)), float(rng.uniform(0, 4)), 1))
            # every state of the new layer must be reachable, every old one must continue
            for v in nxt:
                if not any(a[1] == v for a in arcs):
                    arcs.append((layers[-1][0], v, 2, 1.0, 1.0, 1))
            for u in layers[-1]:
                if not any(a[0] == u for a in arcs):
                    arcs.append((u, nxt[0], 2, 1.0, 1.0, 1))
            layers.append(nxt)
            t += 1
    finals = {s: (float(rng.uniform(0, 1)), 0.0) for s in layers[-1]}
    arcs.sort(key=lambda a: a[0])
    return klu.make_lattice(key, nstates, arcs, finals)
