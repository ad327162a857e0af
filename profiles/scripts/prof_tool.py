"""Per-kernel CUDA-event times of one tool run on a synthetic batch (diagnostics).
usage: prof_tool.py <tool> <shape> <lattices> [flag=value ...]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
import bench  # noqa: E402

klu = load_package()
tool, shape, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
flags = bench.flags_for(tool)
for kv in sys.argv[4:]:
    k, v = kv.split("=")
    flags[k] = float(v) if "." in v or "e" in v else int(v)
eng = klu.Engine(0)
t0 = time.time()
batch = klu.synth_batch(shape, n, seed=21)
print("gen %.2fs arcs %d" % (time.time() - t0, batch.num_arcs), flush=True)
t0 = time.time()
eng.load(batch)
print("load %.3fs" % (time.time() - t0), eng.load_times(), flush=True)
if os.environ.get("PROF_LOAD"):  # per-kernel times of a second (warm) load + the lazily built indexes
    eng.profile(True)
    eng.load(batch)
    bench.run_tool(eng, klu, tool, flags)
    eng.sync()
    prof = eng.profile_json()
    eng.profile(False)
    print("warm load + first run:", eng.load_times(), flush=True)
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        print("   %-28s %3d launches %10.3f ms" % (k, v["launches"], v["ms"]), flush=True)
for it in range(2):
    eng.profile(True)
    t0 = time.time()
    bench.run_tool(eng, klu, tool, flags)
    eng.sync()
    dt = time.time() - t0
    prof = eng.profile_json()
    eng.profile(False)
    print("run %d: %.3fs wall, %.3g arcs/s" % (it, dt, batch.num_arcs / dt), flush=True)
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        print("   %-28s %3d launches %10.3f ms" % (k, v["launches"], v["ms"]), flush=True)
rows, nbytes = bench.fetch_for(eng, klu, tool)
print("rows", rows, "bytes", nbytes, "load_times", eng.load_times(), flush=True)
