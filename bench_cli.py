#!/usr/bin/env python
"""Whole-binary timing of a drop-in tool, ark parse and write included (SURVEY.md 8d
timing (iii)): writes N synthetic lattices as a binary CompactLattice ark, runs the tool
on it, reports arcs/s of the tool's wall clock.

  python bench_cli.py [--tool lattice-to-word-frame-post] [--shape c2] [--lattices 500]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tool", default="lattice-to-word-frame-post")
ap.add_argument("--shape", default="c2")
ap.add_argument("--lattices", type=int, default=500)
ap.add_argument("--flags", default="--acoustic-scale=0.1")
ap.add_argument("--devices", default=None, help="KLU_DEVICES, e.g. 0,0 for two contexts on GPU 0")
args = ap.parse_args()
klu = load_package()
cfg = klu.lattice.SHAPES[args.shape]
BIN = os.path.join(ROOT, "kaldi-lattice-utils_b200", "bin")
with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
    ark = os.path.join(d, "lat.ark")
    t0 = time.perf_counter()
    subprocess.run([os.path.join(BIN, "klu-synth-lattices")] + [str(cfg[k]) for k in (
        "frames", "states_per_frame", "arcs_per_state", "max_skip", "vocab", "pool_size", "window", "eps_prob",
        "weight_max", "kind")] + [str(0x5EED), str(args.lattices), "ark:" + ark], check=True)
    t_gen = time.perf_counter() - t0
    arcs = int(klu.synth_batch(args.shape, args.lattices, seed=0x5EED).num_arcs)
    env = dict(os.environ)
    if args.devices:
        env["KLU_DEVICES"] = args.devices
    cmd = [os.path.join(BIN, args.tool)] + args.flags.split() + ["ark:" + ark, "ark:" + os.path.join(d, "out.ark")]
    if args.tool == "lattice-char-index-position" or args.tool == "lattice-char-index-segment":
        cmd = [os.path.join(BIN, args.tool)] + args.flags.split() + ["1", "ark:" + ark, "ark:" + os.path.join(d, "out.ark")]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, env=env, capture_output=True)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        sys.stderr.write(r.stderr.decode()[-2000:])
        sys.exit(1)
    if os.environ.get("KLU_TRACE"):
        sys.stderr.write("\n".join(l for l in r.stderr.decode().splitlines() if "time:" in l) + "\n")
    print(json.dumps({"tool": args.tool, "shape": args.shape, "lattices": args.lattices, "arcs": arcs,
                      "ark_bytes": os.path.getsize(ark), "out_bytes": os.path.getsize(os.path.join(d, "out.ark")),
                      "seconds": dt, "arcs_per_s": arcs / dt, "ark_write_seconds": t_gen,
                      "devices": args.devices or "0"}))
