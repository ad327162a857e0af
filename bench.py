#!/usr/bin/env python
"""bench.py -- lattice arcs/sec through the forward-backward + posterior-index
hot path on B200 (metric of BASELINE.json), with roofline and CPU baseline.

  python bench.py --gpus N --steps K --warmup W            (this repo's CUDA path)
  python bench.py --impl reference --gpus N --steps K ...  (reference arm: the CPU
      restatement of the reference's algorithm -- oracle/ -- on all host threads;
      the reference itself needs Kaldi+OpenFst and cannot be built in this image)

A "step" = one pass of the hot path over one batch: BASELINE.json configs[1]
(10k synthetic lattices, ~2k states / ~50k arcs each, 50k vocabulary,
lattice-to-word-frame-post with --acoustic-scale=0.1) PER GPU (weak scaling:
lattices are independent, every rank owns its own shard, no collective on the
data path).  `value` times {alpha sweep, beta sweep, arc-posterior emit, sort by
key, segmented log-add, output ordering} with the packed batch resident in HBM;
`e2e` times klu_load (host packer + H2D from pinned host arrays) + klu_run +
klu_fetch_* (D2H of the whole index) through the C ABI.
"""
import argparse
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

TOOLS = {"frame_post": 3, "segment": 0, "position": 1, "utterance": 2, "fwd_bwd": 7, "prune_dyn_beam": 4,
         "best_path2": 5}
# algorithmic HBM bytes: per arc / per state / per emitted entry
ALGO = {
    # alpha sweep src+g+a (12) + beta sweep dst+g+a (12); alpha+beta written once (16/state)
    "k_log_sweeps": dict(arc=24.0, state=16.0, entry=0.0),
    "k_log_sweeps(fwd)": dict(arc=12.0, state=8.0, entry=0.0),
    # backward sweep + arc posteriors: record read (12 B of it algorithmic) + posterior written (8 B)
    "k_log_sweeps(bwd+post)": dict(arc=20.0, state=16.0, entry=0.0),
    "k_banded_alpha": dict(arc=12.0, state=0.0, entry=0.0, band=8.0),
    "k_emit": dict(arc=20.0, state=0.0, entry=16.0),
    "k_seg_radix_sort": dict(arc=0.0, state=0.0, entry=16.0),
    "k_reduce": dict(arc=0.0, state=0.0, entry=16.0),
    "k_count_scan": dict(arc=8.0, state=0.0, entry=0.0),
    "k_gather": dict(arc=0.0, state=0.0, entry=16.0),
    # arc posterior pre-pass: record (16) + source id (4) read, posterior (8) written
    "k_arc_post": dict(arc=28.0, state=16.0, entry=0.0),
    # frame-synchronous group-by + order: every arc x frame instance reads its arc id (4 B)
    # and the arc's posterior (8 B); every (frame, word, logp) row is written once (12 B)
    # group sums: every arc x frame instance reads its arc id (4 B) and the arc's posterior
    # (8 B); per group 4 B offset read + 4 B log-posterior written
    "k_group_post": dict(arc=0.0, state=0.0, entry=8.0, inst=12.0),
    # per-frame ordering: 4 B logp + 4 B word read, (word, logp) written in order (the frame column is static)
    "k_frame_order": dict(arc=0.0, state=0.0, entry=16.0),
    # fused window kernel: record + source id per arc (20 B), arc id per instance (4 B), offset read +
    # log-posterior written per group (8 B)
    "k_frame_groups": dict(arc=20.0, state=16.0, entry=8.0, inst=4.0),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons DURING the timed region: NVML in-process
    (a few ms per sample); nvidia-smi subprocesses only if NVML is unavailable."""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.stop_flag, self.samples, self.reasons, self.max_mhz = gpu, False, [], set(), None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(gpu)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def sample_nvml(self):
        n = self.nvml
        self.samples.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(
            n, "nvmlDeviceGetCurrentClocksEventReasons") else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20),
                          ("hw_thermal_slowdown", 0x40)):
            if r & bit:
                self.reasons.add(name)

    def sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5)
        f = [x.strip() for x in out.stdout.strip().split(",")]
        self.samples.append(float(f[0]))
        self.max_mhz = float(f[1])
        for n, v in zip(names, f[2:]):
            if v.lower().startswith("active"):
                self.reasons.add(n)

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.sample_nvml()
                else:
                    self.sample_smi()
            except Exception:
                pass
            time.sleep(0.005 if self.nvml is not None else 0.2)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def dist_setup(n):
    """Returns (rank, world, reduce_max, barrier)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world == 1:
        return 0, 1, (lambda x: x), (lambda: None)
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    return rank, world, reduce_max, barrier


def flags_for(tool):
    if tool == "prune_dyn_beam":  # BASELINE.json configs[2]
        return dict(max_arcs=20000, max_states=1500, beam_ratio=0.9)
    return dict(acoustic_scale=0.1)


def fetch_for(eng, klu, tool, out=None):
    t = TOOLS[tool]
    if t == klu.FRAME_POST:
        r = eng.fetch_frame_post(out=out)
        return int(r[0][-1]), sum(int(x.nbytes) for x in r)
    if t == klu.SEGMENT:
        r = eng.fetch_segment()
        return int(r[0][-1]), sum(int(x.nbytes) for x in r)
    if t == klu.POSITION:
        r = eng.fetch_position()
        return int(r[0][-1]), sum(int(x.nbytes) for x in r)
    if t == klu.UTTERANCE:
        r = eng.fetch_utterance()
        return int(r[0][-1]), sum(int(x.nbytes) for x in r)
    if t == klu.FWD_BWD:
        r = eng.fetch_fwd_bwd()
        return 0, sum(int(x.nbytes) for x in r)
    if t == klu.PRUNE_DYN_BEAM:
        r = eng.fetch_prune()
        return int(r[0][-1]), sum(int(x.nbytes) for x in r)
    if t == klu.BEST_PATH2:
        r = eng.fetch_best_path2()
        return int(r[0][-1]), sum(int(x.nbytes) for x in r)
    raise ValueError(tool)


def run_reference(args, rank, world):
    """Reference arm: CPU restatement on all host threads, bounded sample."""
    if rank != 0:
        return
    klu = load_package()
    from oracle import ora
    ora.build()
    cores = os.cpu_count() or 1
    n = args.ref_lattices
    batch = klu.synth_batch(args.shape, n, seed=args.seed)
    lats = batch.lattices()
    tool = {"frame_post": ora.FRAME_POST, "segment": ora.SEGMENT, "position": ora.POSITION,
            "utterance": ora.UTTERANCE}[args.tool]
    for _ in range(min(args.warmup, 1)):
        ora.run_batch(tool, lats[:cores], cores, **flags_for(args.tool))
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        ora.run_batch(tool, lats, cores, **flags_for(args.tool))
        times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    v = batch.num_arcs / sec
    sample = "%d lattices of the workload (%d arcs) per step, %d host threads" % (n, batch.num_arcs, cores)
    print(json.dumps({
        "impl": "reference", "metric": "lattice arcs/sec (fwd-bwd + word-position index)", "value": v,
        "unit": "arcs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sec, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, args.lattices),
        "cpu_baseline": {"value": v, "unit": "arcs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "arcs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def workload_config(args, nlat):
    return {"workload": "BASELINE.json configs[1]: lattice-to-word-frame-post forward-backward on %d synthetic "
                        "lattices per GPU (~2k states, ~50k arcs, 50k vocab), --acoustic-scale=0.1" % nlat
            if args.tool == "frame_post" and args.shape == "c2" else
            "%s on %d synthetic '%s' lattices per GPU" % (args.tool, nlat, args.shape),
            "tool": args.tool, "shape": args.shape, "lattices_per_gpu": nlat, "seed": args.seed,
            "l2_policy": "inputs (packed arcs, >10 GB) far larger than the 126 MB L2; no explicit flush",
            "parallelism": "independent lattice shards per GPU, no collective"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tool", default="frame_post", choices=sorted(TOOLS))
    ap.add_argument("--shape", default="c2")
    ap.add_argument("--lattices", type=int, default=10000, help="lattices per GPU")
    ap.add_argument("--ref-lattices", type=int, default=2500, help="bounded CPU sample (lattices) per step")
    ap.add_argument("--seed", type=int, default=0x5EED)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-slices", type=int, default=8, help="slices of the pipelined end-to-end run (1 = off)")
    ap.add_argument("--e2e-contexts", type=int, default=3, help="contexts (host threads) of the pipelined run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        run_reference(args, rank, int(os.environ.get("WORLD_SIZE", "1")))
        return

    rank, world, reduce_max, barrier = dist_setup(args.gpus)
    klu = load_package()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    eng = klu.Engine(local)
    tool = TOOLS[args.tool]
    flags = flags_for(args.tool)
    nlat = args.lattices

    # ---- synthetic shard of this rank, generated straight into pinned host memory
    keep = []

    def alloc(nbytes):
        buf = eng.pinned(nbytes)
        keep.append(buf)
        return buf

    t0 = time.perf_counter()
    batch = klu.synth_batch(args.shape, nlat, seed=args.seed, first_id=rank * nlat, alloc=alloc)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    eng.load(batch)
    t_load = time.perf_counter() - t0
    st = eng.stats()
    arcs, states = st["arcs"], st["states"]

    # ---- device-resident timing (value) ----
    for _ in range(max(args.warmup, 3)):
        eng.run(tool, **flags)
    eng.sync()
    launches0 = eng.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    eng.sync()
    eng.timer_start()
    for _ in range(args.steps):
        eng.run(tool, **flags)
    ms = eng.timer_stop()
    barrier()
    sampler.stop_flag = True
    launches = eng.launch_count() - launches0
    ms = reduce_max(ms)
    ms_per_step = ms / args.steps
    entries, _ = fetch_for(eng, klu, args.tool)
    value = world * arcs / (ms_per_step * 1e-3)

    # ---- per-kernel profile (CUDA events on the launching stream) ----
    eng.profile(True)
    for _ in range(2):
        eng.run(tool, **flags)
    prof = eng.profile_json()
    eng.profile(False)
    tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
    dom = max(prof, key=lambda k: prof[k]["ms"])
    peak, peak_src = peaks()
    band = st["band"]

    def algo_bytes(name):
        a = ALGO.get(name, dict(arc=0, state=0, entry=0))
        return (a["arc"] * arcs + a["state"] * states + a["entry"] * entries + a.get("band", 0) * band +
                a.get("inst", 0) * st.get("frame_instances", 0))

    kern = {}
    for k, v in prof.items():
        per_launch_ms = v["ms"] / v["launches"]
        ab = algo_bytes(k)
        kern[k] = {"launches_per_step": v["launches"] / 2, "ms_per_launch": per_launch_ms,
                   "share": v["ms"] / tot_ms,
                   "algo_GBps": ab / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else None}
    dom_ms = prof[dom]["ms"] / prof[dom]["launches"]
    achieved = algo_bytes(dom) / (dom_ms * 1e-3) / 1e9
    # whole-pipeline algorithmic model of SURVEY.md 8d: 76 B/arc + 28 B/state
    pipe_bytes = 76.0 * arcs + 28.0 * states
    traffic = None  # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if dom in tj:
            traffic = tj[dom]["dram_bytes_per_arc"] * arcs
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes(dom),
                "pipeline_frac_76B_per_arc": pipe_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                "kernels": kern}

    # ---- end to end through the C ABI: host arrays -> index on the host ----
    # (a) one klu_load + klu_run + klu_fetch call sequence over the whole shard;
    # (b) the same calls pipelined the way the drop-in tools do it (KLU_DEVICES): the shard
    #     cut into slices by arc count, a few contexts on this GPU taking slices in turn, one
    #     host thread each, so that the upload of one slice overlaps packing, run and result
    #     download of the previous one.  `e2e.value` is (b); (a) is reported beside it.
    # the caller hands over per-state arc counts instead of a per-arc source array
    # (klu_lattices.state_num_arcs; OpenFst stores arcs grouped by state)
    narcs = batch.state_num_arcs()
    narcs_pinned = eng.pinned_array(np.int32, max(narcs.size, 1))
    narcs_pinned[:narcs.size] = narcs
    narcs = narcs_pinned[:narcs.size]
    h2d = sum(int(x.nbytes) for x in (batch.state_off, batch.arc_off, narcs, batch.dst, batch.label, batch.dur,
                                      batch.graph, batch.acoustic, batch.fin_graph, batch.fin_acoustic,
                                      batch.fin_dur))
    d2h = 0
    out = None
    if args.tool == "frame_post":  # results land in pinned host buffers
        out = (eng.pinned_array(np.int32, entries), eng.pinned_array(np.int32, entries),
               eng.pinned_array(np.float32, entries))
    e2e_ms, parts = [], []
    for i in range(args.e2e_steps + 1):
        barrier()
        t0 = time.perf_counter()
        eng.load(batch, state_num_arcs=narcs)
        t1 = time.perf_counter()
        eng.run(tool, **flags)
        eng.sync()
        t2 = time.perf_counter()
        _, d2h = fetch_for(eng, klu, args.tool, out)
        eng.sync()
        dt = time.perf_counter() - t0
        if i > 0:
            e2e_ms.append(1e3 * dt)
            parts.append((1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t0 + dt - t2)))
    single_step = reduce_max(sum(e2e_ms) / len(e2e_ms)) if e2e_ms else None
    pipe_step = None
    if args.e2e_steps > 0 and args.e2e_slices > 1 and args.tool == "frame_post":
        eng.close()  # its device memory goes to the pipeline contexts
        nsl = min(args.e2e_slices, nlat)
        cuts = np.searchsorted(batch.arc_off, np.linspace(0, batch.arc_off[-1], nsl + 1)[1:-1]).tolist()
        cuts = [0] + [int(x) for x in cuts] + [nlat]
        subs = [batch.slice(a, b) for a, b in zip(cuts[:-1], cuts[1:])]
        sub_narcs = [narcs[int(batch.state_off[a]):int(batch.state_off[b])] for a, b in zip(cuts[:-1], cuts[1:])]
        engines = [klu.Engine(local) for _ in range(min(args.e2e_contexts, nsl))]
        row_off = np.concatenate([[0], np.cumsum([0] * nsl)])  # filled by the first pass
        sub_rows = [0] * nsl

        def work(k, record):
            e = engines[k]
            for j in range(k, nsl, len(engines)):
                e.load(subs[j], state_num_arcs=sub_narcs[j])
                e.run(tool, **flags)
                if record:
                    sub_rows[j] = int(e.offsets()[-1])
                else:
                    a = int(row_off[j])
                    o = tuple(x[a:a + sub_rows[j]] for x in out)
                    e.fetch_frame_post(out=o)

        pipe_ms = []
        for i in range(args.e2e_steps + 2):
            barrier()
            t0 = time.perf_counter()
            th = [threading.Thread(target=work, args=(k, i == 0)) for k in range(len(engines))]
            for t in th:
                t.start()
            for t in th:
                t.join()
            dt = time.perf_counter() - t0
            if i == 0:
                row_off = np.concatenate([[0], np.cumsum(sub_rows)])
                assert int(row_off[-1]) == entries
            elif i > 1:
                pipe_ms.append(1e3 * dt)
        pipe_step = reduce_max(sum(pipe_ms) / len(pipe_ms))
        for e in engines:
            e.close()
    e2e_step = pipe_step if pipe_step else single_step
    e2e = {"value": world * arcs / (e2e_step * 1e-3) if e2e_step else None, "unit": "arcs/s",
           "ms_per_step": e2e_step, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "single_call_ms_per_step": single_step,
           "single_call_ms_load_run_fetch": [round(float(np.mean([p[k] for p in parts])), 2) for k in range(3)]
           if parts else None,
           "pipeline": {"slices": args.e2e_slices, "contexts": args.e2e_contexts} if pipe_step else None,
           "note": "klu_load (H2D of the caller's pinned SoA arrays + device packer) + klu_run + klu_fetch "
                   "(D2H of the full index into pinned buffers), wall clock, max over ranks; value = the "
                   "pipelined call sequence when `pipeline` is set, else the single call sequence"}

    # ---- CPU baseline (rank 0, bounded sample of the same workload) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.tool in ("frame_post", "segment", "position",
                                                                                "utterance"):
        from oracle import ora
        ora.build()
        cores = os.cpu_count() or 1
        n = min(4 * args.ref_lattices, nlat)  # ~10-30 s of CPU work at full size
        sub = batch.slice(0, n)
        otool = {"frame_post": ora.FRAME_POST, "segment": ora.SEGMENT, "position": ora.POSITION,
                 "utterance": ora.UTTERANCE}[args.tool]
        t0 = time.perf_counter()
        ora.run_batch(otool, sub.lattices(), cores, **flags)
        dt = time.perf_counter() - t0
        cpu = {"value": sub.num_arcs / dt, "unit": "arcs/s", "cores": cores, "kind": "port",
               "sample": "first %d lattices of the workload (%d arcs), one oracle run on %d host threads, %.1f s"
                         % (n, sub.num_arcs, cores, dt)}

    if rank == 0:
        print(json.dumps({
            "metric": "lattice arcs/sec (fwd-bwd + word-position index)", "value": value, "unit": "arcs/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, nlat),
            "batch": {"lattices": st["lattices"], "states": states, "arcs": arcs, "levels": st["levels"],
                      "index_entries": entries, "gen_s": t_gen, "load_s": t_load},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": sampler.summary()}))
    if eng.h:
        eng.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
