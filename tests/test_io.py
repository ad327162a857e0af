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
