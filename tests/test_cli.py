"""The drop-in command-line tools (kaldi-lattice-utils_b200/bin/*): flags, usage
and exit codes on CPU; README worked examples and table formats on the GPU."""
import os
import struct
import subprocess

import pytest

from util import GOLD, goldens

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "kaldi-lattice-utils_b200", "bin")
TOOLS = ["lattice-word-index-segment", "lattice-word-index-position", "lattice-word-index-utterance",
         "lattice-to-word-frame-post", "lattice-prune-dyn-beam", "lattice-best-path2", "lattice-char-index-position",
         "lattice-to-word-position-post", "lattice-char-index-segment",
         "lattice-to-transcript-length-dist", "lattice-prune-arcs"]


def run(tool, *args, env=None, stdin=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([os.path.join(BIN, tool)] + list(args), capture_output=True, env=e, input=stdin)


@pytest.mark.parametrize("tool", TOOLS)
def test_help_and_usage_exit_codes(tool):
    r = run(tool, "--help")
    assert r.returncode == 0 and b"Usage:" in r.stderr and b"--acoustic-scale" in r.stderr
    r = run(tool)  # wrong number of positional arguments -> usage + exit(1)
    assert r.returncode == 1 and b"Usage:" in r.stderr


def test_tool_specific_flags_are_registered():
    assert b"--beam" in run("lattice-word-index-position", "--help").stderr
    assert b"--include-words" in run("lattice-word-index-segment", "--help").stderr
    assert b"--rho-label" in run("lattice-word-index-utterance", "--help").stderr
    h = run("lattice-prune-dyn-beam", "--help").stderr
    for f in (b"--beam-ratio", b"--min-beam", b"--max-arcs", b"--max-states"):
        assert f in h
    h = run("lattice-char-index-position", "--help").stderr
    for f in (b"--nbest", b"--other-groups", b"--determinize-delta"):
        assert f in h
    assert b"--beam " not in run("lattice-to-word-frame-post", "--help").stderr


def test_invalid_option_is_an_error():
    r = run("lattice-word-index-segment", "--no-such-flag=1", "ark:a", "ark:b")
    assert r.returncode in (255, 1) and b"Invalid option" in r.stderr


def test_tools_fail_loudly_without_a_gpu():
    """No CPU fallback in the product path: on a box without a usable CUDA device the tools
    stop with the reference's error exit code and say why."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = run("lattice-word-index-position", "ark:" + os.path.join(GOLD, "lattice.ark.txt"), "ark,t:-")
    assert r.returncode == 255 and b"GPU engine" in r.stderr and r.stdout == b""
    r = run("lattice-char-index-position", "1", "ark:" + os.path.join(GOLD, "lattice.char.ark.txt"), "ark,t:-")
    assert r.returncode == 1 and b"GPU engine" in r.stderr and r.stdout == b""


# ------------------------------------------------------------------ GPU ---------
WORD = "ark:" + os.path.join(GOLD, "lattice.ark.txt")
CHAR = "ark:" + os.path.join(GOLD, "lattice.char.ark.txt")


@pytest.mark.gpu
@pytest.mark.parametrize("tool,key", [("lattice-word-index-utterance", "utterance"),
                                      ("lattice-word-index-segment", "segment"),
                                      ("lattice-word-index-position", "position")])
def test_readme_examples_exact_stdout(tool, key):
    r = run(tool, WORD, "ark,t:-")
    assert r.returncode == 0, r.stderr.decode()
    assert r.stdout.decode().rstrip("\n").rstrip(" ") == goldens()[key]
    # BasicTupleVectorHolder text form: fields space-terminated, '; ' separators
    assert r.stdout.decode().endswith(" \n")


@pytest.mark.gpu
def test_binary_tuple_table(tmp_path):
    out = tmp_path / "seg.ark"
    r = run("lattice-word-index-segment", WORD, "ark:" + str(out))
    assert r.returncode == 0, r.stderr.decode()
    b = out.read_bytes()
    assert b.startswith(b"lat1 \0B")
    p = len(b"lat1 \0B")
    assert b[p] == 4
    n = struct.unpack("<i", b[p + 1:p + 5])[0]
    assert n == 10
    p += 5
    rows = []
    for _ in range(n):
        vals = []
        for _k in range(3):
            assert b[p] == 4
            vals.append(struct.unpack("<i", b[p + 1:p + 5])[0])
            p += 5
        assert b[p] == 8
        vals.append(struct.unpack("<d", b[p + 1:p + 9])[0])
        p += 9
        rows.append(tuple(vals))
    assert p == len(b)
    assert rows[0][:3] == (2, 12, 16) and rows[0][3] == 0.0


@pytest.mark.gpu
def test_stdin_pipe_scp_and_batching(tmp_path):
    text = open(os.path.join(GOLD, "lattice.ark.txt"), "rb").read()
    two = text + b"\n" + text.replace(b"lat1", b"lat2")
    r = run("lattice-word-index-segment", "ark:-", "ark,t:-", stdin=two, env={"KLU_BATCH_ARCS": "1"})
    assert r.returncode == 0, r.stderr.decode()
    lines = r.stdout.decode().strip("\n").split("\n")
    assert [x.split()[0] for x in lines] == ["lat1", "lat2"]          # input order kept across batches
    assert lines[0].split(" ", 1)[1] == lines[1].split(" ", 1)[1]
    r2 = run("lattice-word-index-segment", "ark:cat %s |" % os.path.join(GOLD, "lattice.ark.txt"), "ark,t:-")
    assert r2.stdout.decode() == lines[0] + "\n"
    scp = tmp_path / "l.scp"
    scp.write_text("lat1 %s:4\n" % os.path.join(GOLD, "lattice.ark.txt"))   # offset just after the key token
    r3 = run("lattice-word-index-segment", "scp:" + str(scp), "ark,t:-")
    assert r3.returncode == 0, r3.stderr.decode()
    assert r3.stdout.decode() == lines[0] + "\n"


@pytest.mark.gpu
@pytest.mark.parametrize("tool", ["lattice-word-index-position", "lattice-to-word-frame-post"])
def test_multi_context_waves_keep_input_order(tool):
    # KLU_DEVICES: one worker thread + context per listed GPU (here the same GPU
    # twice), one lattice per batch, three waves: output must equal the single-context run
    text = open(os.path.join(GOLD, "lattice.ark.txt"), "rb").read()
    five = b"\n".join(text.replace(b"lat1", b"lat%d" % i) for i in range(1, 6))
    one = run(tool, "ark:-", "ark,t:-", stdin=five)
    multi = run(tool, "ark:-", "ark,t:-", stdin=five, env={"KLU_DEVICES": "0,0", "KLU_BATCH_ARCS": "1"})
    assert one.returncode == 0 and multi.returncode == 0, multi.stderr.decode()
    assert multi.stdout == one.stdout
    assert [x.split()[0] for x in multi.stdout.decode().strip("\n").split("\n")] == ["lat%d" % i for i in range(1, 6)]


@pytest.mark.gpu
def test_length_dist_text():
    r = run("lattice-to-transcript-length-dist", WORD, "ark,t:-")
    assert r.returncode == 0, r.stderr.decode()
    assert r.stdout.decode() == "lat1 [ 7 0 ] \n"


@pytest.mark.gpu
def test_position_post_text():
    r = run("lattice-to-word-position-post", WORD, "ark,t:-")
    assert r.returncode == 0, r.stderr.decode()
    # Posterior text form: one [ word logp ... ] block per transcript position
    assert r.stdout.decode() == ("lat1 [ 2 -0.2231435 1 -1.609438 ] [ 3 -0.2231435 4 -1.609438 ] [ 5 0 ] [ 2 0 ] "
                                 "[ 6 0 ] [ 7 0 ] [ 8 0 ] \n")


@pytest.mark.gpu
def test_frame_post_and_best_path_text():
    r = run("lattice-to-word-frame-post", WORD, "ark,t:-")
    assert r.returncode == 0, r.stderr.decode()
    out = r.stdout.decode()
    assert out.startswith("lat1 [ 2 -0.2231435 1 -1.609438 ] [ 2 -0.2231435 1 -1.609438 ] [ 2 -0.2231435 4 -1.609438 ]")
    assert out.count("[") == 33                                       # one bracket per frame
    r = run("lattice-best-path2", WORD, "ark,t:-")
    assert r.returncode == 0, r.stderr.decode()
    assert r.stdout.decode() == "lat1 2 3 5 2 6 7 8 \n"
    assert b"best cost is 0.4 over 33 frames" in r.stderr


@pytest.mark.gpu
def test_prune_dyn_beam_lattice_roundtrip(tmp_path):
    # no limits: lattice comes back unchanged (text), and the binary form can be
    # read back by another tool with identical results
    r = run("lattice-prune-dyn-beam", WORD, "ark,t:-")
    assert r.returncode == 0, r.stderr.decode()
    assert b"was not pruned" in r.stderr
    lines = r.stdout.decode().split("\n")
    assert lines[0].strip() == "lat1"
    assert lines[1].split("\t")[:3] == ["0", "1", "1"] and lines[1].split("\t")[3] == "1.609438,0,1_28"
    binout = tmp_path / "p.lats"
    r = run("lattice-prune-dyn-beam", "--max-arcs=7", WORD, "ark:" + str(binout))
    assert r.returncode == 0, r.stderr.decode()
    assert b"pruned #states from 10 to 8 and #arcs from 10 to 7" in r.stderr
    r2 = run("lattice-word-index-segment", "ark:" + str(binout), "ark,t:-")
    assert r2.returncode == 0, r2.stderr.decode()
    assert r2.stdout.decode() == ("lat1 2 0 4 0 ; 2 12 16 0 ; 3 4 8 0 ; 5 8 12 0 ; 6 16 22 0 ; 7 22 27 0 ; "
                                  "8 27 33 0 \n")


@pytest.mark.gpu
def test_error_path_exit_code(tmp_path):
    bad = tmp_path / "bad.txt"
    bad.write_text("cyc\n0 1 1 0,0,1\n1 0 2 0,0,1\n1\n\n")
    r = run("lattice-word-index-segment", "ark:" + str(bad), "ark,t:-")
    assert r.returncode == 255 and b"cyclic" in r.stderr           # main() returns -1


@pytest.mark.gpu
def test_prune_arcs_lattice_roundtrip(tmp_path):
    # default --beam=inf: everything is put back (the lattice comes out unchanged up to the order
    # of a state's arcs); a tiny beam drops the most probable arcs, as the reference's code does
    r = run("lattice-prune-arcs", WORD, "ark,t:-")
    assert r.returncode == 0, r.stderr.decode()
    assert b"pruned #states from 10 to 10 and #arcs from 10 to 10" in r.stderr
    assert r.stdout.decode().split("\n")[0].strip() == "lat1"
    r = run("lattice-prune-arcs", "--beam=0.5", WORD, "ark,t:-")
    assert r.returncode == 0, r.stderr.decode()
    assert b"#arcs from 10 to" in r.stderr
    r = run("lattice-prune-arcs", "--beam=-1", WORD, "ark,t:-")
    assert r.returncode == 255 and b"must be in the open range" in r.stderr
