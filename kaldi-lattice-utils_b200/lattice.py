"""Lattice containers, Kaldi text-table parsing/formatting and the synthetic
workload generator (host side, numpy + ctypes; no CUDA here).

Text formats follow the reference's I/O contract: CompactLattice
text entries as in kwsbin2/egs/lattice.ark.txt, the 5-column Lattice form as in
kwsbin2/egs/lattice.char.ark.txt, and BasicTupleVectorHolder text output
(util/basic-tuple-vector-holder.h:149-168).
"""
import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
INF = np.float32(np.inf)


@dataclass
class Lattice:
    key: str
    nstates: int
    src: np.ndarray
    dst: np.ndarray
    label: np.ndarray
    dur: np.ndarray
    graph: np.ndarray
    acoustic: np.ndarray
    fin_graph: np.ndarray
    fin_acoustic: np.ndarray
    fin_dur: np.ndarray

    @property
    def narcs(self):
        return int(self.src.size)


@dataclass
class LatticeBatch:
    """Concatenated SoA arrays of many lattices (the klu_lattices layout)."""
    keys: list
    state_off: np.ndarray
    arc_off: np.ndarray
    src: np.ndarray
    dst: np.ndarray
    label: np.ndarray
    dur: np.ndarray
    graph: np.ndarray
    acoustic: np.ndarray
    fin_graph: np.ndarray
    fin_acoustic: np.ndarray
    fin_dur: np.ndarray

    def __len__(self):
        return len(self.state_off) - 1

    @property
    def num_states(self):
        return int(self.state_off[-1])

    @property
    def num_arcs(self):
        return int(self.arc_off[-1])

    def __getitem__(self, i):
        s0, s1 = int(self.state_off[i]), int(self.state_off[i + 1])
        e0, e1 = int(self.arc_off[i]), int(self.arc_off[i + 1])
        return Lattice(self.keys[i], s1 - s0, self.src[e0:e1], self.dst[e0:e1], self.label[e0:e1], self.dur[e0:e1],
                       self.graph[e0:e1], self.acoustic[e0:e1], self.fin_graph[s0:s1], self.fin_acoustic[s0:s1],
                       self.fin_dur[s0:s1])

    def lattices(self):
        return [self[i] for i in range(len(self))]

    def state_num_arcs(self):
        """Arcs leaving each state (int32, concatenated over the batch): the optional
        klu_lattices.state_num_arcs input that replaces the per-arc source array."""
        out = np.zeros(int(self.state_off[-1]), np.int32)
        for l in range(len(self)):
            s0, s1 = int(self.state_off[l]), int(self.state_off[l + 1])
            e0, e1 = int(self.arc_off[l]), int(self.arc_off[l + 1])
            if e1 > e0:
                out[s0:s1] = np.bincount(self.src[e0:e1], minlength=s1 - s0)
        return out

    def slice(self, lo, hi):
        s0, s1 = int(self.state_off[lo]), int(self.state_off[hi])
        e0, e1 = int(self.arc_off[lo]), int(self.arc_off[hi])
        return LatticeBatch(self.keys[lo:hi], self.state_off[lo:hi + 1] - s0, self.arc_off[lo:hi + 1] - e0,
                            self.src[e0:e1], self.dst[e0:e1], self.label[e0:e1], self.dur[e0:e1], self.graph[e0:e1],
                            self.acoustic[e0:e1], self.fin_graph[s0:s1], self.fin_acoustic[s0:s1],
                            self.fin_dur[s0:s1])

    @staticmethod
    def from_lattices(lats):
        so = np.zeros(len(lats) + 1, np.int64)
        ao = np.zeros(len(lats) + 1, np.int64)
        for i, l in enumerate(lats):
            so[i + 1] = so[i] + l.nstates
            ao[i + 1] = ao[i] + l.narcs

        def cat(name, dt):
            if not lats:
                return np.zeros(0, dt)
            return np.ascontiguousarray(np.concatenate([np.asarray(getattr(l, name), dt) for l in lats]))

        return LatticeBatch([l.key for l in lats], so, ao, cat("src", np.int32), cat("dst", np.int32),
                            cat("label", np.int32), cat("dur", np.int32), cat("graph", np.float32),
                            cat("acoustic", np.float32), cat("fin_graph", np.float32), cat("fin_acoustic", np.float32),
                            cat("fin_dur", np.int32))


def make_lattice(key, nstates, arcs, finals):
    """arcs: iterable of (src, dst, label, graph, acoustic, dur), any order (made
    stable-sorted by src); finals: {state: (graph, acoustic[, dur])}."""
    arcs = sorted(arcs, key=lambda a: a[0])
    a = np.array(arcs, dtype=np.float64).reshape(-1, 6)
    fg = np.full(nstates, np.inf, np.float32)
    fa = np.full(nstates, np.inf, np.float32)
    fd = np.zeros(nstates, np.int32)
    for s, w in finals.items():
        fg[s], fa[s] = np.float32(w[0]), np.float32(w[1])
        fd[s] = w[2] if len(w) > 2 else 0
    return Lattice(key, nstates, a[:, 0].astype(np.int32), a[:, 1].astype(np.int32), a[:, 2].astype(np.int32),
                   a[:, 5].astype(np.int32), a[:, 3].astype(np.float32), a[:, 4].astype(np.float32), fg, fa, fd)


def _parse_weight(tok):
    """'g,a,t1_t2_..' (CompactLattice) or 'g,a' (Lattice) -> (g, a, dur)."""
    parts = tok.split(",")
    g = float(parts[0]) if parts[0] != "" else 0.0
    a = float(parts[1]) if len(parts) > 1 and parts[1] != "" else 0.0
    dur = 0
    if len(parts) > 2 and parts[2] != "":
        dur = len(parts[2].split("_"))
    return g, a, dur


def read_text_ark(path_or_lines):
    """Reads a text-mode lattice table (CompactLattice 4-column or Lattice
    5-column entries).  5-column arcs become CompactLattice arcs with label =
    olabel and duration = (ilabel != 0), which is what ConvertLattice [ext]
    yields when no linear chain can be merged (every arc of the reference's
    fixture carries an olabel)."""
    if isinstance(path_or_lines, str):
        with open(path_or_lines) as f:
            lines = f.read().split("\n")
    else:
        lines = list(path_or_lines)
    lats = []
    i = 0
    while i < len(lines):
        if lines[i].strip() == "":
            i += 1
            continue
        key = lines[i].split()[0]
        i += 1
        arcs, finals, maxs = [], {}, -1
        while i < len(lines) and lines[i].strip() != "":
            tok = lines[i].split()
            i += 1
            if len(tok) <= 2:  # final state
                s = int(tok[0])
                w = _parse_weight(tok[1]) if len(tok) == 2 else (0.0, 0.0, 0)
                finals[s] = w
                maxs = max(maxs, s)
            elif len(tok) == 3 or (len(tok) == 4 and "," in tok[3]):
                s, d, lab = int(tok[0]), int(tok[1]), int(tok[2])
                g, a, dur = _parse_weight(tok[3]) if len(tok) == 4 else (0.0, 0.0, 0)
                arcs.append((s, d, lab, g, a, dur))
                maxs = max(maxs, s, d)
            else:  # Lattice: src dst ilabel olabel [g,a]
                s, d, il, ol = int(tok[0]), int(tok[1]), int(tok[2]), int(tok[3])
                g, a, _ = _parse_weight(tok[4]) if len(tok) == 5 else (0.0, 0.0, 0)
                arcs.append((s, d, ol, g, a, 1 if il != 0 else 0))
                maxs = max(maxs, s, d)
        lats.append(topsort(make_lattice(key, maxs + 1, arcs, finals)))
    return lats


def topsort(lat):
    """TopSortCompactLatticeIfNeeded [ext]: renumber only when some arc has
    src >= dst; OpenFst TopSort order (reverse DFS finishing order, start first)."""
    if lat.narcs == 0 or np.all(lat.src < lat.dst):
        return lat
    n = lat.nstates
    out = [[] for _ in range(n)]
    for e in range(lat.narcs):
        out[int(lat.src[e])].append(e)
    color = [0] * n
    finish = []
    for root in [0] + list(range(n)):
        if color[root]:
            continue
        stack = [(root, 0)]
        color[root] = 1
        while stack:
            s, k = stack.pop()
            if k < len(out[s]):
                stack.append((s, k + 1))
                d = int(lat.dst[out[s][k]])
                if color[d] == 1:
                    raise ValueError("cyclic lattice")
                if color[d] == 0:
                    color[d] = 1
                    stack.append((d, 0))
            else:
                color[s] = 2
                finish.append(s)
    order = np.zeros(n, np.int64)
    for newid, s in enumerate(reversed(finish)):
        order[s] = newid
    arcs = [(int(order[lat.src[e]]), int(order[lat.dst[e]]), int(lat.label[e]), float(lat.graph[e]),
             float(lat.acoustic[e]), int(lat.dur[e])) for e in range(lat.narcs)]
    finals = {int(order[s]): (lat.fin_graph[s], lat.fin_acoustic[s], int(lat.fin_dur[s]))
              for s in range(n) if not (np.isinf(lat.fin_graph[s]) and np.isinf(lat.fin_acoustic[s]))}
    return make_lattice(lat.key, n, arcs, finals)


def kaldi_float(x):
    """Kaldi text streams print floating fields with precision 7 (%.7g)."""
    if x == 0:
        return "0"
    if np.isinf(x):
        return "inf" if x > 0 else "-inf"
    return "%.7g" % x


def format_tuples(key, rows):
    """BasicTupleVectorHolder text form: 'key f1 f2 ; f1 f2 \\n'
    (util/basic-tuple-vector-holder.h:157-166: fields space-terminated, '; '
    between tuples)."""
    parts = []
    for row in rows:
        parts.append("".join((kaldi_float(f) if isinstance(f, float) else str(f)) + " " for f in row))
    return key + " " + "; ".join(parts) + "\n"


# --- synthetic workloads -----------------------------------------------------
class SynthCfg(C.Structure):
    _fields_ = [("kind", C.c_int32), ("frames", C.c_int32), ("states_per_frame", C.c_float),
                ("arcs_per_state", C.c_float), ("max_skip", C.c_int32), ("vocab", C.c_int32),
                ("pool_size", C.c_int32), ("window", C.c_int32), ("eps_prob", C.c_float), ("weight_max", C.c_float)]


# config 2/3 shape (~2k states, ~50k arcs, 50k vocab), config 4
# (deep: ~20k states, ~500k arcs), config 5 (HTR char lattices), and a tiny shape
# for unit tests.
SHAPES = {
    "c2": dict(kind=0, frames=600, states_per_frame=3.3, arcs_per_state=24.0, max_skip=3, vocab=50000, pool_size=200,
               window=20, eps_prob=0.05, weight_max=10.0),
    "c4": dict(kind=0, frames=6000, states_per_frame=3.3, arcs_per_state=24.0, max_skip=3, vocab=50000,
               pool_size=200, window=20, eps_prob=0.05, weight_max=10.0),
    "c5": dict(kind=1, frames=500, states_per_frame=4.0, arcs_per_state=2.0, max_skip=1, vocab=200, pool_size=0,
               window=0, eps_prob=0.0, weight_max=4.0),
    "tiny": dict(kind=0, frames=12, states_per_frame=2.2, arcs_per_state=2.5, max_skip=3, vocab=12, pool_size=6,
                 window=6, eps_prob=0.1, weight_max=4.0),
    "small": dict(kind=0, frames=60, states_per_frame=3.0, arcs_per_state=8.0, max_skip=3, vocab=300, pool_size=30,
                  window=15, eps_prob=0.05, weight_max=10.0),
    "tinychar": dict(kind=1, frames=14, states_per_frame=2.0, arcs_per_state=1.5, max_skip=1, vocab=6, pool_size=0,
                     window=0, eps_prob=0.0, weight_max=3.0),
}

_HOSTLIB = None


def hostlib():
    global _HOSTLIB
    if _HOSTLIB is None:
        path = os.path.join(_HERE, "libklu_host.so")
        if not os.path.exists(path):
            raise RuntimeError("libklu_host.so is not built: run __graft_entry__.build() or `make -C %s`" % _HERE)
        L = C.CDLL(path)
        L.klu_synth_sizes.argtypes = [C.POINTER(SynthCfg), C.c_uint64, C.c_uint64, C.c_int32, C.c_void_p, C.c_void_p,
                                      C.c_int]
        L.klu_synth_fill.argtypes = [C.POINTER(SynthCfg), C.c_uint64, C.c_uint64, C.c_int32] + [C.c_void_p] * 11 + [
            C.c_int]
        _HOSTLIB = L
    return _HOSTLIB


def synth_batch(shape, n, seed=0x5EED, first_id=0, nthreads=None, alloc=None, **overrides):
    """Generates lattices first_id .. first_id+n-1 of a named shape.  `alloc`
    (optional) is a callable (nbytes) -> writable buffer, e.g. pinned memory."""
    cfgd = dict(SHAPES[shape] if isinstance(shape, str) else shape)
    cfgd.update(overrides)
    cfg = SynthCfg(**cfgd)
    L = hostlib()
    nthreads = nthreads or min(os.cpu_count() or 1, 64)
    so = np.zeros(n + 1, np.int64)
    ao = np.zeros(n + 1, np.int64)
    L.klu_synth_sizes(C.byref(cfg), seed, first_id, n, so.ctypes.data, ao.ctypes.data, nthreads)
    S, E = int(so[-1]), int(ao[-1])

    def arr(count, dt):
        if alloc is None:
            return np.zeros(count, dt)
        return np.frombuffer(alloc(max(count, 1) * np.dtype(dt).itemsize), dtype=dt, count=count)

    src, dst, label, dur = (arr(E, np.int32) for _ in range(4))
    g, a = arr(E, np.float32), arr(E, np.float32)
    fg, fa, fd = arr(S, np.float32), arr(S, np.float32), arr(S, np.int32)
    rc = L.klu_synth_fill(C.byref(cfg), seed, first_id, n, so.ctypes.data, ao.ctypes.data, src.ctypes.data,
                          dst.ctypes.data, label.ctypes.data, dur.ctypes.data, g.ctypes.data, a.ctypes.data,
                          fg.ctypes.data, fa.ctypes.data, fd.ctypes.data, nthreads)
    if rc != 0:
        raise RuntimeError("klu_synth_fill: size mismatch")
    keys = ["utt%07d" % (first_id + i) for i in range(n)]
    return LatticeBatch(keys, so, ao, src, dst, label, dur, g, a, fg, fa, fd)
