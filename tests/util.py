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
