"""Table I/O layer of the drop-in tools (host/kaldi_io.{h,cc}) through klu-copy-lattices:
text <-> binary round trips, the byte layout Kaldi/OpenFst readers expect ("\\0B" marker,
VectorFst header with the arc count left at zero), the threaded block reader against the
sequential one, and error exits.  No GPU."""
import os
import struct
import subprocess

import pytest

from util import GOLD

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "kaldi-lattice-utils_b200", "bin")
COPY = os.path.join(BIN, "klu-copy-lattices")
SYNTH = os.path.join(BIN, "klu-synth-lattices")
WORD = os.path.join(GOLD, "lattice.ark.txt")
FST_MAGIC = 2125659606


def copy(*args, threads=None):
    env = dict(os.environ)
    if threads is not None:
        env["KLU_IO_THREADS"] = str(threads)
    return subprocess.run([COPY] + list(args), capture_output=True, env=env)


def synth_ark(path, n, seed=7):
    import sys
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package
    cfg = load_package().lattice.SHAPES["small"]
    subprocess.run([SYNTH] + [str(cfg[k]) for k in (
        "frames", "states_per_frame", "arcs_per_state", "max_skip", "vocab", "pool_size", "window", "eps_prob",
        "weight_max", "kind")] + [str(seed), str(n), "ark:" + path], check=True)


def test_text_binary_text_round_trip(tmp_path):
    b, t1, t2 = (str(tmp_path / x) for x in ("a.bin", "a.txt", "b.txt"))
    assert copy("ark:" + WORD, "ark:" + b).returncode == 0
    assert copy("ark:" + WORD, "ark,t:" + t1).returncode == 0
    assert copy("ark:" + b, "ark,t:" + t2).returncode == 0
    assert open(t1, "rb").read() == open(t2, "rb").read()
    assert open(t1).read().startswith("lat1")


def test_binary_layout_is_what_kaldi_and_openfst_write(tmp_path):
    b = str(tmp_path / "a.bin")
    assert copy("ark:" + WORD, "ark:" + b).returncode == 0
    d = open(b, "rb").read()
    key, rest = d.split(b" ", 1)
    assert key == b"lat1"
    assert rest[:2] == b"\0B"  # InitKaldiOutputStream
    p = 2
    assert struct.unpack_from("<i", rest, p)[0] == FST_MAGIC
    p += 4
    n = struct.unpack_from("<i", rest, p)[0]
    assert rest[p + 4:p + 4 + n] == b"vector"
    p += 4 + n
    n = struct.unpack_from("<i", rest, p)[0]
    assert rest[p + 4:p + 4 + n] == b"compactlattice44"
    p += 4 + n
    version, flags = struct.unpack_from("<ii", rest, p)
    p += 8 + 8  # + properties
    start, nstates, narcs = struct.unpack_from("<qqq", rest, p)
    assert (version, flags & 3, start) == (2, 0, 0) and nstates > 0
    assert narcs == 0  # VectorFst::WriteFst never sets it; readers must count


def test_reader_does_not_trust_the_header_arc_count(tmp_path):
    """A wrong (non-zero) arc count in the header must not matter either."""
    b, t1, t2 = (str(tmp_path / x) for x in ("a.bin", "a.txt", "b.txt"))
    assert copy("ark:" + WORD, "ark:" + b).returncode == 0
    d = bytearray(open(b, "rb").read())
    off = d.index(b"compactlattice44") + 16 + 4 + 4 + 8 + 8 + 8
    assert struct.unpack_from("<q", d, off)[0] == 0
    struct.pack_into("<q", d, off, 12345)
    open(b, "wb").write(d)
    assert copy("ark:" + b, "ark,t:" + t1).returncode == 0
    assert copy("ark:" + WORD, "ark,t:" + t2).returncode == 0
    assert open(t1, "rb").read() == open(t2, "rb").read()


@pytest.mark.parametrize("threads", [1, 4])
def test_block_reader_equals_sequential_reader(tmp_path, threads):
    ark, a, b = (str(tmp_path / x) for x in ("in.ark", "par.ark", "seq.ark"))
    synth_ark(ark, 40)
    assert copy("ark:" + ark, "ark:" + a, threads=threads).returncode == 0
    r = copy("--sequential", "ark:" + ark, "ark:" + b)
    assert r.returncode == 0 and b"Copied 40 lattices" in r.stderr
    assert open(a, "rb").read() == open(b, "rb").read()
    # and the copy parses back to itself
    c = str(tmp_path / "again.ark")
    assert copy("ark:" + a, "ark:" + c).returncode == 0
    assert open(a, "rb").read() == open(c, "rb").read()


def test_pipe_and_stdin_inputs(tmp_path):
    b, t1, t2 = (str(tmp_path / x) for x in ("a.bin", "a.txt", "b.txt"))
    assert copy("ark:" + WORD, "ark:" + b).returncode == 0
    assert copy("ark:cat %s |" % b, "ark,t:" + t1).returncode == 0
    assert copy("ark:" + b, "ark,t:" + t2).returncode == 0
    assert open(t1, "rb").read() == open(t2, "rb").read()


@pytest.mark.parametrize("mode", [[], ["--sequential"]])
def test_truncated_archive_is_an_error(tmp_path, mode):
    ark, cut = str(tmp_path / "in.ark"), str(tmp_path / "cut.ark")
    synth_ark(ark, 6)
    d = open(ark, "rb").read()
    open(cut, "wb").write(d[:len(d) - 1000])
    r = copy(*mode, "ark:" + cut, "ark:" + str(tmp_path / "out.ark"))
    assert r.returncode == 1 and b"ERROR" in r.stderr


def test_text_field_formatting_matches_iostream_and_printf():
    """Text tables are formatted with std::to_chars; the binary checks itself against
    `os << int` / printf("%.7g") on six million values (every exponent, subnormals,
    infinities, NaN, typical log-posteriors)."""
    r = subprocess.run([COPY, "--format-selftest"], capture_output=True)
    assert r.returncode == 0 and b"identical" in r.stderr, r.stderr.decode()


# ---- lattices that are not topologically sorted, and corrupt state ids -----------------
def _parse_binary_lattice(d):
    """(key, states) of a one-entry binary CompactLattice archive; states = list of
    (final (g, a, tids) or None, [(label, g, a, tids, dst)])."""
    key, rest = d.split(b" ", 1)
    assert rest[:2] == b"\0B"
    p = 2 + 4
    for _ in range(2):
        n = struct.unpack_from("<i", rest, p)[0]
        p += 4 + n
    p += 4 + 4 + 8
    start, nstates, _ = struct.unpack_from("<qqq", rest, p)
    p += 24
    assert start == 0
    states = []
    for _ in range(nstates):
        g, a, sz = struct.unpack_from("<ffi", rest, p)
        p += 12
        tids = struct.unpack_from("<%di" % sz, rest, p)
        p += 4 * sz
        na = struct.unpack_from("<q", rest, p)[0]
        p += 8
        arcs = []
        for _ in range(na):
            il, ol, ag, aa, asz = struct.unpack_from("<iiffi", rest, p)
            p += 20
            at = struct.unpack_from("<%di" % asz, rest, p)
            p += 4 * asz
            dst = struct.unpack_from("<i", rest, p)[0]
            p += 4
            arcs.append((ol, ag, aa, at, dst))
        states.append(((g, a, tids) if g != float("inf") else None, arcs))
    assert p == len(rest)
    return key, states


def _write_binary_lattice(key, states):
    out = key + b" \0B" + struct.pack("<i", FST_MAGIC)
    for name in (b"vector", b"compactlattice44"):
        out += struct.pack("<i", len(name)) + name
    out += struct.pack("<iiQqqq", 2, 0, 0, 0, len(states), 0)
    inf = float("inf")
    for fin, arcs in states:
        g, a, tids = fin if fin else (inf, inf, ())
        out += struct.pack("<ffi", g, a, len(tids)) + struct.pack("<%di" % len(tids), *tids)
        out += struct.pack("<q", len(arcs))
        for ol, ag, aa, at, dst in arcs:
            out += struct.pack("<iiffi", ol, ol, ag, aa, len(at)) + struct.pack("<%di" % len(at), *at)
            out += struct.pack("<i", dst)
    return out


def _permuted(states, perm):
    """State i becomes perm[i] (perm[0] == 0: the start stays 0)."""
    out = [None] * len(states)
    for i, (fin, arcs) in enumerate(states):
        out[perm[i]] = (fin, [(ol, g, a, t, perm[d]) for ol, g, a, t, d in arcs])
    return out


def _strip_tids(text):
    rows = []
    for line in text.decode().splitlines():
        f = line.split()
        if f and "," in f[-1]:
            w = f[-1].split(",")
            f[-1] = "%s,%s,%d" % (w[0], w[1], len([x for x in w[2].split("_") if x]) if len(w) > 2 else 0)
        rows.append(" ".join(f))
    return rows


@pytest.mark.parametrize("mode", [[], ["--sequential"]])
@pytest.mark.parametrize("tids", [[], ["--no-tids"]])
def test_unsorted_binary_lattice_is_topsorted_on_read(tmp_path, mode, tids):
    """A binary CompactLattice whose state ids are not in topological order goes through
    TopSortCompactLatticeIfNeeded [ext] on read -- with and without the transition-id strings
    (every tool but lattice-prune-dyn-beam reads without them)."""
    b, u, t1, t2 = (str(tmp_path / x) for x in ("a.bin", "u.bin", "a.txt", "u.txt"))
    assert copy("ark:" + WORD, "ark:" + b).returncode == 0
    key, states = _parse_binary_lattice(open(b, "rb").read())
    n = len(states)
    perm = [0] + list(range(n - 1, 0, -1))  # reverse everything but the start
    open(u, "wb").write(_write_binary_lattice(key, _permuted(states, perm)) * 3)
    r = copy(*mode, *tids, "ark:" + u, "ark,t:" + t2)
    assert r.returncode == 0, r.stderr.decode()
    assert copy("ark:" + b, "ark,t:" + t1).returncode == 0
    want = _strip_tids(open(t1, "rb").read())
    got = _strip_tids(open(t2, "rb").read())
    # the README lattice has one topological order up to the two parallel branches; the sorted
    # copy must describe the same arcs (src < dst everywhere) with the same weights and durations
    assert len(got) == 3 * len(want)
    one = got[:len(want)]
    assert one[0] == want[0] == "lat1"
    arcs = [r.split() for r in one[1:] if len(r.split()) == 4]
    assert all(int(a[0]) < int(a[1]) for a in arcs)
    assert sorted((a[2], a[3]) for a in arcs) == sorted(
        (a[2], a[3]) for a in (r.split() for r in want[1:]) if len(a) == 4)


@pytest.mark.parametrize("bad_dst", [-1, 10 ** 6])
@pytest.mark.parametrize("tids", [[], ["--no-tids"]])
def test_out_of_range_state_ids_are_an_error(tmp_path, bad_dst, tids):
    b, u = str(tmp_path / "a.bin"), str(tmp_path / "u.bin")
    assert copy("ark:" + WORD, "ark:" + b).returncode == 0
    key, states = _parse_binary_lattice(open(b, "rb").read())
    fin, arcs = states[2]
    states[2] = (fin, [arcs[0][:4] + (bad_dst,)] + arcs[1:])
    open(u, "wb").write(_write_binary_lattice(key, states))
    for mode in ([], ["--sequential"]):
        r = copy(*mode, *tids, "ark:" + u, "ark,t:" + str(tmp_path / "o.txt"))
        assert r.returncode == 1 and b"ERROR" in r.stderr and b"outside" in r.stderr


def test_negative_state_id_in_text_is_an_error(tmp_path):
    t = str(tmp_path / "neg.txt")
    open(t, "w").write("bad\n0 1 5 1.0,2.0,3_4\n1 -2 6 1.0,2.0,7\n1\n\n")
    r = copy("ark:" + t, "ark,t:" + str(tmp_path / "o.txt"))
    assert r.returncode == 1 and b"negative state id" in r.stderr


# ---- pipes are streamed and their exit status counts; write failures are errors ----------
def test_failing_input_pipe_is_an_error(tmp_path):
    o = str(tmp_path / "o.txt")
    r = copy("ark:cat /nonexistent/file 2>/dev/null |", "ark,t:" + o)
    assert r.returncode == 1 and b"nonzero return status" in r.stderr
    # output produced, then a failure: still an error (a truncated table must not pass)
    r = copy("ark:cat %s; exit 3 |" % WORD, "ark,t:" + o)
    assert r.returncode == 1 and b"nonzero return status" in r.stderr


def test_failing_output_pipe_is_an_error(tmp_path):
    r = copy("ark:" + WORD, "ark,t:| cat > /dev/null; exit 3")
    assert r.returncode == 1 and b"nonzero return status" in r.stderr
    ok = str(tmp_path / "ok.txt")
    r = copy("ark:" + WORD, "ark,t:| cat > " + ok)
    assert r.returncode == 0 and open(ok).read().startswith("lat1")


def test_write_failure_is_an_error():
    if not os.path.exists("/dev/full"):
        pytest.skip("no /dev/full")
    r = copy("ark:" + WORD, "ark,t:/dev/full")
    assert r.returncode == 1 and b"ERROR" in r.stderr


def test_large_pipe_is_streamed(tmp_path):
    """Pipe input and output go through a fixed-size buffer (no whole-archive copy in RAM):
    the copy of a pipe equals the copy of the file."""
    ark, a, b = (str(tmp_path / x) for x in ("in.ark", "a.ark", "b.ark"))
    synth_ark(ark, 60)
    assert copy("ark:cat %s |" % ark, "ark:| cat > " + a).returncode == 0
    assert copy("ark:" + ark, "ark:" + b).returncode == 0
    assert open(a, "rb").read() == open(b, "rb").read()
