"""Per-kernel CUDA-event times of lattice-best-path2 on lattices pruned by lattice-prune-dyn-beam
(BASELINE.json configs[2]).  usage: prof_bp2.py <lattices>"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
import bench  # noqa: E402

klu = load_package()
n = int(sys.argv[1])
eng = klu.Engine(0)
batch = klu.synth_batch("c2", n, seed=21)
eng.load(batch)
bench.run_tool(eng, klu, "prune_dyn_beam", bench.flags_for("prune_dyn_beam"))
pruned = eng.pruned_batch()
eng.load(pruned)
print("pruned arcs", pruned.num_arcs, flush=True)
for it in range(3):
    eng.profile(True)
    t0 = time.time()
    bench.run_tool(eng, klu, "best_path2", {})
    eng.sync()
    dt = time.time() - t0
    prof = eng.profile_json()
    eng.profile(False)
    print("run %d: %.3fs wall" % (it, dt), flush=True)
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        print("   %-28s %3d launches %10.3f ms" % (k, v["launches"], v["ms"]), flush=True)
