"""ctypes face of the CPU oracle (oracle/klu_oracle.cc).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product never imports it.
"""
import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SEGMENT, POSITION, UTTERANCE, FRAME_POST, PRUNE_DYN_BEAM, BEST_PATH2, CHAR_POSITION = range(7)
POSITION_POST = 8
CHAR_SEGMENT = 9
LENGTH_DIST = 14
FWD_BWD = 18
PRUNE_ARCS, BRUTE_PRUNE_ARCS = 19, 20
BRUTE_SEGMENT, BRUTE_POSITION, BRUTE_FRAME, BRUTE_UTTERANCE = 10, 11, 12, 13
TOP_ORDER, BRUTE_BEST_PATH2, BRUTE_PRUNE = 15, 16, 17

INT_MAX = 2**31 - 1


class OraLat(C.Structure):
    _fields_ = [("nstates", C.c_int32), ("narcs", C.c_int32),
                ("src", C.c_void_p), ("dst", C.c_void_p), ("label", C.c_void_p), ("dur", C.c_void_p),
                ("graph", C.c_void_p), ("acoustic", C.c_void_p),
                ("fin_graph", C.c_void_p), ("fin_acoustic", C.c_void_p), ("fin_dur", C.c_void_p)]


class OraOpts(C.Structure):
    _fields_ = [("acoustic_scale", C.c_float), ("graph_scale", C.c_float), ("insertion_penalty", C.c_float),
                ("beam", C.c_float), ("beam_ratio", C.c_float), ("min_beam", C.c_float),
                ("max_arcs", C.c_int32), ("max_states", C.c_int32), ("nbest", C.c_int32),
                ("include_words", C.c_void_p), ("n_include", C.c_int32),
                ("exclude_words", C.c_void_p), ("n_exclude", C.c_int32),
                ("group_labels", C.c_void_p), ("group_ids", C.c_void_p), ("n_group_labels", C.c_int32),
                ("inc_groups", C.c_void_p), ("n_inc_groups", C.c_int32),
                ("del_groups", C.c_void_p), ("n_del_groups", C.c_int32)]


def build():
    """Compile the oracle with its Makefile (idempotent)."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libklu_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.ora_run.restype = C.c_void_p
        L.ora_run.argtypes = [C.c_int, C.POINTER(OraLat), C.POINTER(OraOpts)]
        L.ora_run_batch.restype = C.c_int64
        L.ora_run_batch.argtypes = [C.c_int, C.POINTER(OraLat), C.c_int64, C.POINTER(OraOpts), C.c_int]
        L.ora_error.restype = C.c_char_p
        L.ora_error.argtypes = [C.c_void_p]
        L.ora_nrows.restype = C.c_int64
        L.ora_nrows.argtypes = [C.c_void_p, C.c_int]
        L.ora_get_i.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ora_get_d.argtypes = [C.c_void_p, C.c_void_p]
        L.ora_get_f.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ora_get_str.restype = C.c_char_p
        L.ora_get_str.argtypes = [C.c_void_p, C.c_int64]
        L.ora_scalar_i.restype = C.c_int64
        L.ora_scalar_i.argtypes = [C.c_void_p, C.c_int]
        L.ora_scalar_d.restype = C.c_double
        L.ora_scalar_d.argtypes = [C.c_void_p, C.c_int]
        L.ora_free.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def _c32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def make_lat(lat):
    """lat: any object with nstates, src, dst, label, dur, graph, acoustic,
    fin_graph, fin_acoustic, fin_dur.  Returns (OraLat, keepalive)."""
    keep = [_c32(lat.src), _c32(lat.dst), _c32(lat.label), _c32(lat.dur), _f32(lat.graph), _f32(lat.acoustic),
            _f32(lat.fin_graph), _f32(lat.fin_acoustic), _c32(lat.fin_dur)]
    o = OraLat(int(lat.nstates), int(keep[0].size), *[_ptr(k) for k in keep])
    return o, keep


def make_opts(acoustic_scale=1.0, graph_scale=1.0, insertion_penalty=0.0, beam=float("inf"), beam_ratio=0.9,
              min_beam=1e-3, max_arcs=INT_MAX, max_states=INT_MAX, nbest=100, include_words=(), exclude_words=(),
              label_group=None, inc_groups=(), del_groups=()):
    keep = [_c32(list(include_words)), _c32(list(exclude_words))]
    label_group = label_group or {}
    keep += [_c32(list(label_group.keys())), _c32(list(label_group.values())), _c32(list(inc_groups)),
             _c32(list(del_groups))]
    o = OraOpts(acoustic_scale, graph_scale, insertion_penalty, beam, beam_ratio, min_beam, max_arcs, max_states,
                nbest, _ptr(keep[0]), keep[0].size, _ptr(keep[1]), keep[1].size, _ptr(keep[2]), _ptr(keep[3]),
                keep[2].size, _ptr(keep[4]), keep[4].size, _ptr(keep[5]), keep[5].size)
    return o, keep


@dataclass
class OraResult:
    i: list = field(default_factory=list)   # up to four int32 columns
    d: np.ndarray = None
    f: list = field(default_factory=list)   # up to two float32 columns
    s: list = field(default_factory=list)
    s0: int = 0
    s1: int = 0
    ds0: float = 0.0
    ds1: float = 0.0


def run(tool, lat, **opts):
    L = lib()
    ol, k1 = make_lat(lat)
    oo, k2 = make_opts(**opts)
    h = L.ora_run(tool, C.byref(ol), C.byref(oo))
    try:
        err = L.ora_error(h).decode()
        if err:
            raise RuntimeError("oracle: " + err)
        r = OraResult()
        for c in range(4):
            n = L.ora_nrows(h, c)
            a = np.zeros(n, np.int32)
            L.ora_get_i(h, c, _ptr(a))
            r.i.append(a)
        n = L.ora_nrows(h, 4)
        r.d = np.zeros(n, np.float64)
        L.ora_get_d(h, _ptr(r.d))
        for c in range(2):
            n = L.ora_nrows(h, 5 + c)
            a = np.zeros(n, np.float32)
            L.ora_get_f(h, c, _ptr(a))
            r.f.append(a)
        r.s = [L.ora_get_str(h, i).decode() for i in range(L.ora_nrows(h, 7))]
        r.s0, r.s1 = L.ora_scalar_i(h, 0), L.ora_scalar_i(h, 1)
        r.ds0, r.ds1 = L.ora_scalar_d(h, 0), L.ora_scalar_d(h, 1)
        return r
    finally:
        L.ora_free(h)


def run_batch(tool, lats, nthreads, **opts):
    """Times nothing itself; runs `tool` over all lattices on nthreads host
    threads and returns the total number of output rows."""
    L = lib()
    arr = (OraLat * len(lats))()
    keep = []
    for i, lat in enumerate(lats):
        o, k = make_lat(lat)
        arr[i] = o
        keep.append(k)
    oo, k2 = make_opts(**opts)
    return L.ora_run_batch(tool, arr, len(lats), C.byref(oo), nthreads)


# --- convenience views -------------------------------------------------------
def segment(lat, **o):
    r = run(SEGMENT, lat, **o)
    return list(zip(r.i[0].tolist(), r.i[1].tolist(), r.i[2].tolist(), r.d.tolist()))


def position(lat, **o):
    r = run(POSITION, lat, **o)
    return list(zip(r.i[0].tolist(), r.i[1].tolist(), r.i[2].tolist(), r.i[3].tolist(), r.d.tolist()))


def utterance(lat, **o):
    r = run(UTTERANCE, lat, **o)
    return list(zip(r.i[0].tolist(), r.d.tolist()))


def frame_post(lat, **o):
    """Returns list (one per frame) of lists of (word, float32 logp)."""
    r = run(FRAME_POST, lat, **o)
    frames = [[] for _ in range(r.s0)]
    for k, w, p in zip(r.i[0].tolist(), r.i[1].tolist(), r.f[0].tolist()):
        frames[k].append((w, p))
    return frames


def position_post(lat, **o):
    """latbin/lattice-to-word-position-post: list (one per transcript position) of lists of
    (word, float32 logp)."""
    r = run(POSITION_POST, lat, **o)
    pos = [[] for _ in range(r.s0)]
    for k, w, p in zip(r.i[0].tolist(), r.i[1].tolist(), r.f[0].tolist()):
        pos[k].append((w, p))
    return pos


def length_dist(lat, **o):
    """latbin/lattice-to-transcript-length-dist: [(length, float32 logp)] in output order."""
    r = run(LENGTH_DIST, lat, **o)
    return list(zip(r.i[0].tolist(), r.f[0].tolist()))


def best_path2(lat, **o):
    r = run(BEST_PATH2, lat, **o)
    return r.i[0].tolist(), r.ds0


def prune_dyn_beam(lat, **o):
    r = run(PRUNE_DYN_BEAM, lat, **o)
    arcs, finals = [], []
    for k in range(len(r.i[0])):
        if r.i[0][k] >= 0:
            arcs.append((int(r.i[0][k]), int(r.i[1][k]), int(r.i[2][k]), int(r.i[3][k]), float(r.f[0][k]),
                         float(r.f[1][k])))
        else:
            finals.append((int(r.i[1][k]), float(r.f[0][k]), float(r.f[1][k])))
    return dict(arcs=arcs, finals=finals, nstates=r.s0, iters=r.s1, beam0=r.ds0, beam=r.ds1)


def _prune_rows(r):
    arcs, finals = [], []
    for k in range(len(r.i[0])):
        if r.i[0][k] >= 0:
            arcs.append((int(r.i[0][k]), int(r.i[1][k]), int(r.i[2][k]), int(r.i[3][k]), float(r.f[0][k]),
                         float(r.f[1][k])))
        else:
            finals.append((int(r.i[1][k]), float(r.f[0][k]), float(r.f[1][k])))
    return arcs, finals


def brute_prune_dyn_beam(lat, **o):
    """lattice-prune-dyn-beam from the list of all paths (tiny lattices): same fields as
    prune_dyn_beam plus `margin`, the distance of the nearest arc to the last cutoff."""
    r = run(BRUTE_PRUNE, lat, **o)
    arcs, finals = _prune_rows(r)
    return dict(arcs=arcs, finals=finals, nstates=r.s0, iters=r.s1, beam0=r.ds0, beam=r.ds1,
                margin=float(r.d[0]) if len(r.d) else float("inf"))


def prune_arcs(lat, **o):
    """latbin/lattice-prune-arcs: the lattice it writes (arcs of a state in the order AddArc left
    them), `first_kept` = index of the first arc put back, `cutoff` = beam - total."""
    r = run(PRUNE_ARCS, lat, **o)
    arcs, finals = _prune_rows(r)
    return dict(arcs=arcs, finals=finals, nstates=r.s0, first_kept=r.s1, cutoff=r.ds0)


def brute_prune_arcs(lat, **o):
    r = run(BRUTE_PRUNE_ARCS, lat, **o)
    arcs, finals = _prune_rows(r)
    return dict(arcs=arcs, finals=finals, nstates=r.s0, first_kept=r.s1, cutoff=r.ds0,
                margin=float(r.d[0]) if len(r.d) else float("inf"))


def brute_best_path2(lat, **o):
    """lattice-best-path2 from the list of all paths: (labels, cost, margin to the next-best
    different label sequence)."""
    r = run(BRUTE_BEST_PATH2, lat, **o)
    return r.i[0].tolist(), r.ds0, r.ds1


def top_order(lat):
    """[ext] fst::TopSort order of a lattice numbered in any way: new id of every old state."""
    return run(TOP_ORDER, lat).i[0].tolist()


def fwd_bwd(lat, **o):
    """Per-state alpha, beta of ComputeLatticeAlphasAndBetas [ext] and its return value."""
    r = run(FWD_BWD, lat, **o)
    n = len(r.d) // 2
    return r.d[:n].copy(), r.d[n:].copy(), r.ds0


def char_position(lat, wspace, other_groups=(), **o):
    label_group = {0: 0}
    for w in wspace:
        label_group[w] = 1
    inc = [INT_MAX]
    for gi, grp in enumerate(other_groups):
        for lab in grp:
            label_group[lab] = gi + 2
        inc.append(gi + 2)
    r = run(CHAR_POSITION, lat, label_group=label_group, inc_groups=inc, del_groups=[1], **o)
    return list(zip(r.s, r.i[0].tolist(), r.i[1].tolist(), r.i[2].tolist(), r.d.tolist()))


def char_segment(lat, wspace, other_groups=(), **o):
    """kwsbin2/lattice-char-index-segment: rows (string, t0, t1, logp)."""
    label_group = {0: 0}
    for w in wspace:
        label_group[w] = 1
    inc = [INT_MAX]
    for gi, grp in enumerate(other_groups):
        for lab in grp:
            label_group[lab] = gi + 2
        inc.append(gi + 2)
    r = run(CHAR_SEGMENT, lat, label_group=label_group, inc_groups=inc, del_groups=[1], **o)
    return list(zip(r.s, r.i[0].tolist(), r.i[1].tolist(), r.d.tolist()))


def brute(tool, lat):
    r = run(tool, lat)
    return {(a, b, c): v for a, b, c, v in zip(r.i[0].tolist(), r.i[1].tolist(), r.i[2].tolist(), r.d.tolist())}
