"""Multi-GPU host logic on CPU: arc-count partitioning, and a world_size-2 gloo run
in which every rank processes its shard (with the CPU oracle standing in for the
engine -- test infrastructure only) and rank 0 reassembles input order."""
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import torch.distributed as dist
from __graft_entry__ import load_package
klu = load_package()
from oracle import ora
ora.build()
dist.init_process_group(backend="gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
lats = klu.synth_batch("small", 11, seed=77).lattices()
def run_shard(ls):
    return [ora.segment(l, acoustic_scale=0.5) for l in ls]
merged = klu.run_sharded(lats, run_shard, rank=rank, world=2)
if rank == 0:
    single = [ora.segment(l, acoustic_scale=0.5) for l in lats]
    assert merged == single, "sharded result differs from the single-process result"
    print("OK", len(merged))
dist.barrier()
dist.destroy_process_group()
'''


def test_partition_balances_and_keeps_input_order(klu):
    rng = np.random.RandomState(0)
    counts = rng.randint(1, 1000, size=200)
    shards = klu.partition_by_arcs(counts, 8)
    assert sorted(i for s in shards for i in s) == list(range(200))
    loads = [int(counts[s].sum()) for s in shards]
    assert max(loads) - min(loads) <= counts.max()
    assert all(s == sorted(s) for s in shards)
    assert klu.partition_by_arcs([5, 3], 1) == [[0, 1]]
    assert klu.partition_by_arcs([], 4) == [[], [], [], []]


def test_gloo_world2_reassembles_input_order(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                              cwd=ROOT) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e.decode()[-2000:]
    assert b"OK 11" in outs[0][0]
