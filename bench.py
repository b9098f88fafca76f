#!/usr/bin/env python
"""bench.py — headline benchmark of the collision hot path (refit -> broadphase -> narrowphase).

Metric (BASELINE.json): candidate+contact pairs/sec (and ms/step) at 1M bodies.  A "step" is one
pass of the path over the scene: axcd_refit + axcd_broadphase + axcd_narrowphase.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

* N = 1: the headline workload — 1,000,000 mixed boxes/spheres, L = 100, seed 3.
* N > 1 (launched by torchrun, one rank per GPU): independent worlds, one 1M-body scene per rank
  (seed 3 + rank), no data-path collective -> weak scaling; `value` = all ranks' pairs / max time.  The
  same invocation then also runs the two multi-GPU configurations BASELINE.json names and reports them
  as extra keys: `c3` (4096 x 256-body worlds split over the ranks, strong scaling, no collective) and
  `c4` (the 16M-body scene in x-slabs with the ghost exchange over NCCL inside libaxcd.so).
* `value`   : inputs already resident in HBM (transforms uploaded once), the fused step as the public API
              runs it (axcd_step_async: one CUDA graph launch), CUDA-event timed per step, L2 flushed
              between timed steps.  The per-stage table comes from separate steps through the staged calls.
* `e2e`     : the same metric through the public API with HOST buffers: every step uploads the
              poses (position + rotation, axcd_set_poses, 28 B per body; the scales are static and resident)
              from pinned host memory (H2D inside the timed region) and reads the contacts back to pinned
              host memory (D2H).  `e2e_full_transforms`: the same with whole 40-byte Transforms every step.
* `--impl reference`: the CPU oracle (the only "reference implementation" that exists for this
              path — the upstream snapshot has no collision code) on all host cores, on the SAME full-size
              workload (one step of the 1M-body scene is a few hundred ms of CPU work).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "axiom-physics-engine_b200"))

METRIC = "candidate+contact pairs/sec at 1M bodies"
UNIT = "pairs/s"
WORKLOADS = {
    "headline": "1M mixed boxes/spheres, L=100, seed 3 (refit + sort + LBVH + GJK/EPA)",
    "C1": "100k mixed boxes/spheres, L=46.4, seed 2",
    "C2": "1M bodies 40% box / 30% sphere / 30% 16-vertex hulls, L=100, seed 4 (EPA-heavy)",
    "C3": "4096 independent 256-body worlds batched, L=6.35, seed 1000+world",
    "C4": "16M-body single scene (scaled by --scale), x-slab decomposition with ghost exchange over NCCL",
}
L2_NOTE = "256 MiB written, then 256 MiB read, between timed steps: L2 cold and clean"
# flops per unit of the FP32-bound kernels: GJK ~2 kflop per pair that runs it, EPA 10-20
# kflop per penetrating pair (midpoint)
GJK_FLOP_PER_PAIR = 2.0e3
EPA_FLOP_PER_PAIR = 15.0e3


def make_config(workload, world, n, npairs, ncon, nepa):
    """The `config` object, identical for our arm and the reference arm of the same workload."""
    return {"workload": WORKLOADS[workload], "bodies_per_gpu": int(n), "candidate_pairs": int(npairs),
            "contacts": int(ncon), "epa_runs": int(nepa), "l2": L2_NOTE,
            "parallelism": "one independent scene per rank, no collective" if world > 1 else "single GPU"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clocks / throttle reasons of one GPU during the timed region (rank 0 only: one sampler
    per job, not one per rank).  NVML through nvidia_ml_py when importable, else nvidia-smi."""

    def __init__(self, device):
        self.device = device
        self.samples = []
        self._stop = threading.Event()
        self._t = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._h = pynvml.nvmlDeviceGetHandleByIndex(device)
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        sm = n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self._h, n.NVML_CLOCK_SM)
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        r = get(self._h)
        bits = (0x8, 0x40, 0x20, 0x4)   # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap
        return [str(sm), str(mx)] + ["Active" if (r & b) else "Not Active" for b in bits]

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}",
                              "--format=csv,noheader,nounits"], capture_output=True,
                             text=True, timeout=5).stdout.strip()
        return [x.strip() for x in out.split(",")] if out else None

    def _run(self):
        while not self._stop.is_set():
            try:
                s = self._sample_nvml() if self._nvml else self._sample_smi()
                if s:
                    self.samples.append(s)
            except Exception:
                pass
            self._stop.wait(0.02 if self._nvml else 0.1)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
                for nme, v in zip(names, s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def oracle_step(O, s, nthreads):
    """One full step of the CPU oracle; returns (pairs, contacts, epa runs, seconds, per-stage seconds)."""
    t0 = time.perf_counter()
    rc, bb = O.refit(s.xf, s.shapes, s.hull, nthreads=nthreads)
    t1 = time.perf_counter()
    pairs = O.broadphase(bb, s.world_id, nthreads=nthreads, cap=max(1024, 8 * s.n))
    t2 = time.perf_counter()
    con, _, nst = O.narrowphase(s.xf, s.shapes, pairs, s.hull, nthreads=nthreads)
    t3 = time.perf_counter()
    return len(pairs), len(con), int(nst.numPenetrating), t3 - t0, (t1 - t0, t2 - t1, t3 - t2)


def run_reference(args):
    """--impl reference: the CPU oracle timed on the host cores (rank 0 only), full-size workload."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import axcd
    import oracle_lib as O
    nthreads = host_threads()
    workload = args.workload if args.workload != "C4" else "headline"
    s = axcd.config_scene(workload)
    # a full step is ~0.3-0.5 s of CPU work on 16 threads: bound the loop so the arm ends within minutes
    warm = max(1, min(args.warmup, 2))
    steps = max(1, min(args.steps, 20))
    for _ in range(warm):
        oracle_step(O, s, nthreads)
    tot_units, tot_s = 0, 0.0
    for _ in range(steps):
        np_, nc, nepa, sec, parts = oracle_step(O, s, nthreads)
        tot_units += np_ + nc
        tot_s += sec
    value = tot_units / tot_s
    sample = (f"{steps} full steps of the same {s.n}-body workload on {nthreads} host threads "
              f"(last step: refit {parts[0]:.2f}s, grid broadphase {parts[1]:.2f}s, narrowphase {parts[2]:.2f}s)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": make_config(workload, world, s.n, np_, nc, nepa),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample,
                         "steps_timed": steps},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the upstream snapshot has no collision code; the reference arm is the in-repo CPU oracle "
                "(oracle/axref.cpp) on the host cores; one scene on rank 0 at every N",
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# measurement helpers (our arm)
# ------------------------------------------------------------------------------------------------------
class Flusher:
    """L2 flush between timed steps: write a 256 MiB buffer (> 126 MB L2), then read another 256 MiB, so the
    cache is cold AND clean when the step starts.  The write alone leaves ~126 MB of the flush buffer dirty in
    L2 and the first kernels of the step pay for its write-back (a cost of the benchmark, not of the path)."""

    def __init__(self, torch, stream):
        self.torch, self.stream = torch, stream
        self.wr = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
        self.rd = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device="cuda")

    def __call__(self, mode="write+read"):
        with self.torch.cuda.stream(self.stream):
            self.wr.fill_(1)
            if mode == "write+read":
                self.rd.sum()


def time_fused_steps(torch, w, stream, flush, steps, warmup):
    """`steps` fused steps (axcd_step_async = one graph launch once stable), each between two CUDA events on
    the context stream with an L2 flush before it.  Returns (sum of ms, last stats)."""
    for _ in range(warmup):
        st = w.step()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(steps):
        flush()
        ev[i][0].record(stream)
        w.step_async()
        ev[i][1].record(stream)
        st = w.stats()           # synchronises; counts
    return sum(a.elapsed_time(b) for a, b in ev), st


STAGE_KEYS = ("refitMs", "sortMs", "buildMs", "pairMs", "pairSortMs", "gjkMs", "epaMs")


def time_staged_steps(w, flush, steps):
    """Per-stage CUDA-event times through the staged calls (direct launches with events between stages)."""
    acc = {k: 0.0 for k in STAGE_KEYS}
    tot = 0.0
    for _ in range(steps):
        flush()
        w.update()               # Broadphase::update(): refit + broadphase
        w.detect_collisions()    # Narrowphase::detectCollisions()
        st = w.stats()
        for k in acc:
            acc[k] += getattr(st, k)
        tot += st.totalMs
    return {k: v / steps for k, v in acc.items()}, tot / steps, st


def stage_table(avg, st, hbm_peak, fp32_peak, generic_pairs):
    """Per-stage roofline rows.  Algorithmic bytes / flops per stage as stated in DESIGN.md section 2."""
    n, npairs, ncon, nepa = st.numBodies, st.numPairs, st.numContacts, st.numPenetrating
    bits_n = max(1, (max(n, 2) - 1).bit_length())          # as bitsFor() in csrc/axcd_api.cu
    key_bits = 3 * max(1, min(10, (bits_n + 2) // 3 + 1))   # Morton bits per axis chosen from N
    passes = (key_bits + 7) // 8
    info = {
        "refitMs": ("refitTmaKernel", "hbm", n * 80),
        "sortMs": ("mortonKernel (keys + digit histograms) + onesweep radix sort (%d passes)" % passes, "hbm", n * 32 + n * 16 * passes),
        "buildMs": ("leaf gather fused with the range tree + Karras topology/fit (32-byte nodes)", "hbm", n * (24 + 32) + n * 32 + n * 32),
        "pairMs": ("findPairsDenseKernel (LBVH traversal, dense leaf tests)", "hbm", n * 32 + npairs * 8),
        "pairSortMs": ("pair counting sort (scan + scatter + block-cooperative segment sort)", "hbm", n * 12 + npairs * (8 + 4 + 4 + 8)),
        "epaMs": ("epaKernel (+fallback)", "fp32", nepa * EPA_FLOP_PER_PAIR),
    }
    if generic_pairs > 0.05 * max(1, npairs):
        info["gjkMs"] = ("classify + closedFormKernel + gjkKernel + slotKernel", "fp32", generic_pairs * GJK_FLOP_PER_PAIR)
    else:
        info["gjkMs"] = ("narrowClosedFusedKernel (classify + sphere / box closed forms incl. box-box SAT + in-order compaction)", "hbm",
                         npairs * (8 + 2 + 2 * 56) + ncon * 40)
    rows = []
    for k, ms in avg.items():
        name, bound, work = info[k]
        row = {"stage": k[:-2], "kernels": name, "ms": round(ms, 4), "bound": bound}
        if bound == "hbm":
            gbs = work / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            row.update({"algorithmic_bytes": int(work), "achieved_gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / hbm_peak, 4)})
        else:
            tf = work / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
            row.update({"algorithmic_flops": int(work), "achieved_tflops": round(tf, 3),
                        "frac_of_fp32_peak": round(tf / fp32_peak, 4) if fp32_peak else None})
        rows.append(row)
    return rows, info


def roofline_of(rows, info, avg, hbm_peak, peak_src, fp32_peak, workload, sm_count=None, sm_mhz=None):
    dom = max(avg, key=avg.get)
    name, bound, work = info[dom]
    row = next(r for r in rows if r["stage"] == dom[:-2])
    if bound == "hbm":
        ach, peak, unit = row["achieved_gbs"], hbm_peak, "GB/s"
        src = peak_src
    else:
        ach, peak, unit = row["achieved_tflops"], fp32_peak, "TFLOP/s"
        src = "measured in this run: FMA-chain kernel, 8 chains per thread, 2048 threads per SM (axcd_test_fp32_peak)"
    # DRAM traffic of the dominant kernel per launch from this round's `ncu --set full` capture of the same
    # workload — used only while the captured kernel time still matches the live one (else null: stale)
    traffic = None
    issue = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        ent = tj.get(workload, {}).get(dom[:-2])
        if ent and abs(ent["ncu_kernel_ms"] - avg[dom]) <= 0.25 * avg[dom]:
            traffic = ent["dram_bytes_per_launch"]
            if ent.get("warp_instructions_per_launch") and sm_count and sm_mhz:
                # what actually bounds a kernel that is neither streaming nor FP32-dense: issue slots.  Warp instructions
                # per launch (a property of kernel + scene, from the same ncu capture) over the live launch time, against
                # 4 schedulers per SM at the SM clock sampled during the run.
                ginst = ent["warp_instructions_per_launch"] / (avg[dom] * 1e-3) / 1e9
                peak_i = sm_count * 4 * sm_mhz * 1e-3
                issue = {"warp_instructions_per_launch": ent["warp_instructions_per_launch"],
                         "achieved_ginst_per_s": round(ginst, 1), "peak_ginst_per_s": round(peak_i, 1),
                         "frac": round(ginst / peak_i, 4), "active_lanes_per_instruction": ent.get("threads_per_instruction"),
                         "lsu_data_pipe_pct_of_peak_ncu": ent.get("lsu_data_pipe_pct_of_peak"),
                         "source": "instruction count and lanes: profiles/r02_traffic.json (ncu --set full of this kernel); time and clock: this run"}
    except Exception:
        pass
    return {"kernel": name, "bound": bound, "achieved": ach, "peak": round(peak, 2) if peak else None, "unit": unit,
            "frac": round(ach / peak, 4) if peak else None, "traffic": traffic, "peak_source": src,
            "launch_ms": round(avg[dom], 4), "issue": issue,
            "note": "dominant stage of the step by CUDA-event time through the staged calls; algorithmic work per "
                    "DESIGN.md section 2; every stage's own row is in `stages`.  `issue` (when the ncu capture matches "
                    "this run) is the kernel's share of the SMs' instruction-issue slots, with the LSU data-pipe share ncu "
                    "measured beside it: the LBVH walk is co-limited by those two, not by HBM"}


def measure_next_rows(w, s, stream, hbm_peak, device):
    """Timings of the stages built on top of the hot path on the bench scene.
    Reported beside the headline, never inside it."""
    import numpy as np
    import torch
    import axcd
    out = {}
    reps = 5
    ms = 0.0
    w.step()
    w.build_manifolds()   # first use allocates the manifold buffer: keep it out of the timing
    w.stats()
    for _ in range(reps):
        w.step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        w.build_manifolds()
        e1.record(stream)
        st = w.stats()
        ms += e0.elapsed_time(e1)
    ms /= reps
    nbytes = st.numContacts * (40 + 88)
    out["manifolds"] = {"ms": round(ms, 4), "contacts": int(st.numContacts), "contact_points": int(st.contactPointCount),
                        "algorithmic_bytes": int(nbytes), "achieved_gbs": round(nbytes / (ms * 1e-3) / 1e9, 1),
                        "frac_of_hbm_peak": round(nbytes / (ms * 1e-3) / 1e9 / hbm_peak, 4)}
    rng = np.random.default_rng(0)
    nq = 1 << 18
    L = float(s.xf[:, :3].max())
    o = rng.uniform(0, L, (nq, 3)).astype(np.float32)
    d = rng.normal(size=(nq, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros(nq, axcd.RAY_DT)
    rays["ox"], rays["oy"], rays["oz"] = o[:, 0], o[:, 1], o[:, 2]
    rays["dx"], rays["dy"], rays["dz"] = d[:, 0], d[:, 1], d[:, 2]
    rays["tMax"] = 50.0
    boxes = np.concatenate([o - 1.0, o + 1.0], axis=1)
    w.raycast(rays[:1024])
    ray_s = 1e9
    for _ in range(3):           # blocking host calls: best of three
        t0 = time.perf_counter()
        hits = w.raycast(rays)
        ray_s = min(ray_s, time.perf_counter() - t0)
    t1 = time.perf_counter()
    qh = w.query_aabbs(boxes)   # first call sizes the output (601 + retry is part of the public call)
    t2 = time.perf_counter()
    q_s = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        qh = w.query_aabbs(boxes)
        q_s = min(q_s, time.perf_counter() - t0)
    out["raycast"] = {"rays": nq, "t_max": 50.0, "hit_fraction": round(float((hits["body"] != axcd.NO_HIT).mean()), 4),
                      "mrays_per_s_e2e": round(nq / ray_s / 1e6, 2)}
    out["aabb_query"] = {"queries": nq, "hits": int(len(qh)), "mqueries_per_s_e2e": round(nq / q_s / 1e6, 2),
                         "first_call_ms": round(1e3 * (t2 - t1), 3)}
    wc = axcd.CollisionWorld.for_scene(s, device=device, stream=stream.cuda_stream, aabbMargin=0.05,
                                       flags=axcd.FLAG_TEMPORAL_COHERENCE, pairs_per_body=12)
    full = wc.step()
    skip_ms, sk = 0.0, None
    for _ in range(reps):
        wc.set_transforms(s.xf)
        sk = wc.step()
        skip_ms += sk.totalMs
    out["temporal_coherence"] = {"margin": 0.05, "full_step_ms": round(full.totalMs, 4), "full_step_pairs": int(full.numPairs),
                                 "cached_step_ms": round(skip_ms / reps, 4), "broadphase_skipped": int(sk.broadphaseSkipped),
                                 "moved_bodies": int(sk.movedBodies), "contacts": int(sk.numContacts)}
    wc.close()
    return out


def generic_pairs_of(scene, st):
    """Pairs that run GJK: those with a hull or capsule on either side (estimated from the shape mix: the
    bench scenes place shapes independently of position)."""
    import numpy as np
    t = scene.shapes["type"]
    frac = float(np.mean((t == 2) | (t == 4)))
    return st.numPairs * (1.0 - (1.0 - frac) ** 2)


def measure_side_workload(torch, axcd, scene, flags, stream, flush, device, hbm_peak, peak_src, fp32_peak, workload,
                          steps, generic_pairs_fn):
    """A second workload (or the same one under other flags) measured beside the headline on rank 0."""
    w = axcd.CollisionWorld.for_scene(scene, device=device, stream=stream.cuda_stream, flags=flags)
    tot, st = time_fused_steps(torch, w, stream, flush, steps, 3)
    avg, staged_total, st2 = time_staged_steps(w, flush, min(steps, 10))
    rows, info = stage_table(avg, st2, hbm_peak, fp32_peak, int(generic_pairs_fn(scene, st)))
    out = {"workload": WORKLOADS[workload], "flags": int(flags), "ms_per_step": round(tot / steps, 4),
           "pairs_per_s": (st.numPairs + st.numContacts) / (tot / steps * 1e-3),
           "candidate_pairs": int(st.numPairs), "contacts": int(st.numContacts), "epa_runs": int(st.numPenetrating),
           "graph_launched": int(st.graphLaunched), "ms_per_step_staged_calls": round(staged_total, 4),
           "stages": rows, "roofline": roofline_of(rows, info, avg, hbm_peak, peak_src, fp32_peak, workload)}
    w.close()
    return out


# ------------------------------------------------------------------------------------------------------
# multi-GPU configurations (N > 1)
# ------------------------------------------------------------------------------------------------------
def allreduce(torch, dist, vals, op):
    t = torch.tensor(vals, dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=op)
    return [float(x) for x in t]


def measure_c3(torch, dist, axcd, rank, world, local, stream, flush, steps):
    """Config C3: 4096 x 256-body worlds, worlds [r*W/R, (r+1)*W/R) on rank r, no collective on the data path."""
    from axcd import sharding
    s, _ = sharding.shard_worlds(axcd.config_scene("C3"), rank, world)
    w = axcd.CollisionWorld.for_scene(s, device=local, stream=stream.cuda_stream)
    for _ in range(3):
        w.step()
    dist.barrier()
    torch.cuda.synchronize()
    tot, st = time_fused_steps(torch, w, stream, flush, steps, 0)
    dist.barrier()
    torch.cuda.synchronize()
    units = float((st.numPairs + st.numContacts) * steps)
    tmax = allreduce(torch, dist, [tot], dist.ReduceOp.MAX)[0]
    usum, psum, csum = allreduce(torch, dist, [units, float(st.numPairs), float(st.numContacts)], dist.ReduceOp.SUM)
    w.close()
    return {"workload": WORKLOADS["C3"], "scaling": "strong", "worlds_per_gpu": int(s.num_worlds),
            "bodies_per_gpu": int(st.numBodies), "ms_per_step": tmax / steps, "pairs_per_s": usum / (tmax * 1e-3),
            "candidate_pairs": int(psum), "contacts": int(csum), "steps": steps, "graph_launched": int(st.graphLaunched),
            "timing": "CUDA events around the fused step on each rank, L2 flushed, max over ranks"}


def measure_c4(torch, dist, axcd, rank, world, local, scale, steps):
    """Config C4: one 16M-body scene in x-slabs, one slab per rank; the ghost exchange runs inside libaxcd.so over
    NCCL (axcd_slab_step).  Wall-clock per step (barrier + synchronize on both sides, max over ranks) with the
    device times of the exchange and of the collision step beside it."""
    import numpy as np
    from axcd import sharding
    s = axcd.config_scene("C4", scale=scale)
    edges = sharding.plan_slabs(s.xf[:, 0], world)
    mine = np.nonzero(sharding.owner_of(s.xf[:, 0], edges) == rank)[0]
    owned = sharding._subset(s, mine)
    rk = sharding.SlabRank(owned, mine.astype(np.uint32), edges, rank, world, device=local)
    del s
    rk.init_native(dist)
    for _ in range(3):
        st = rk.step_native()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev_ms = xch_ms = 0.0
    for _ in range(steps):
        st = rk.step_native()
        dev_ms += st.totalMs
        xch_ms += st.exchangeMs
    dist.barrier()
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    smax, dmax, xmax = allreduce(torch, dist, [sec, dev_ms, xch_ms], dist.ReduceOp.MAX)
    psum, csum, gsum = allreduce(torch, dist, [float(st.numPairs), float(st.numContacts), float(st.ghostBodies)],
                                 dist.ReduceOp.SUM)
    out = {"workload": WORKLOADS["C4"], "scaling": "strong", "scale": scale, "bodies_total": int(round(16_000_000 * scale)),
           "bodies_per_gpu_owned": int(owned.n), "ghost_bodies_total": int(gsum),
           "ms_per_step_wall": 1e3 * smax / steps, "ms_per_step_device_collision": dmax / steps,
           "ms_per_step_device_exchange": xmax / steps,
           "pairs_per_s": (psum + csum) * steps / smax, "candidate_pairs": int(psum), "contacts": int(csum),
           "steps": steps, "graph_launched": int(st.graphLaunched),
           "parallelism": f"{world} x-slabs; ghost records selected on the device, sizes by ncclAllGather, records by "
                          "grouped ncclSend/ncclRecv inside libaxcd.so; x* ownership rule in the traversal kernel",
           "timing": "wall clock over the timed steps between barrier+synchronize pairs, max over ranks; device times "
                     "are CUDA events on each rank's stream (max over ranks)"}
    rk.close()
    return out


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import axcd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the collision path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL logs (its version banner too) off stdout: one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.Stream()
    flush = Flusher(torch, stream)
    hbm_peak, peak_src = load_peaks()

    if args.workload == "C4":   # stand-alone slab run (the default N>1 run reports it as the `c4` key)
        if world < 2:
            raise SystemExit("--workload C4 needs torchrun with at least 2 ranks")
        c4 = measure_c4(torch, dist, axcd, rank, world, local, args.scale, max(1, min(args.steps, 20)))
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": c4["pairs_per_s"], "unit": UNIT, "n_gpus": world,
                              "steps": c4["steps"], "warmup": 3, "ms_per_step": c4["ms_per_step_wall"],
                              "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                              "data": "synthetic", "config": {"workload": WORKLOADS["C4"]}, "c4": c4, "gpu_launches": None}))
        dist.destroy_process_group()
        return

    # ---- workload ---------------------------------------------------------------------------------
    scaling = "weak"
    if args.workload == "headline" and world > 1:
        s = axcd.generate_scene(1_000_000, 3 + rank, 100.0, name="headline")   # one world per rank
    elif args.workload == "C3" and world > 1:
        from axcd import sharding
        s, _ = sharding.shard_worlds(axcd.config_scene("C3"), rank, world)      # worlds [r*W/R, (r+1)*W/R)
        scaling = "strong"
    else:
        s = axcd.config_scene(args.workload)
    w = axcd.CollisionWorld.for_scene(s, device=local, stream=stream.cuda_stream, flags=args.flags)

    # ---- device-resident timing (`value`): the fused step, one CUDA graph launch per step ----------------
    for _ in range(max(args.warmup, 3)):
        st = w.step()
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    total_ms, st = time_fused_steps(torch, w, stream, flush, args.steps, 0)
    barrier()
    clocks = sampler.stop() if sampler else None
    graph_launched = int(st.graphLaunched)
    units = float((st.numPairs + st.numContacts) * args.steps)
    if world > 1:
        total_ms = allreduce(torch, dist, [total_ms], dist.ReduceOp.MAX)[0]
        units = allreduce(torch, dist, [units], dist.ReduceOp.SUM)[0]
    ms_per_step = total_ms / args.steps
    value = units / (total_ms * 1e-3)

    # ---- per-stage times through the staged calls (events between stages, direct launches) ----------------
    stage_steps = max(3, min(args.steps, 20))
    avg, staged_total, st_s = time_staged_steps(w, flush, stage_steps)
    refit_dirty = 0.0
    for _ in range(10):   # the refit stage again under the write-only flush (dirty L2), for the record
        flush("write")
        w.update()
        w.detect_collisions()
        refit_dirty += w.stats().refitMs
    refit_dirty /= 10

    # ---- end-to-end through the public API with host buffers (`e2e`) ---------------------------------
    h_xf = torch.from_numpy(s.xf.copy()).pin_memory()
    h_pose = torch.from_numpy(np.ascontiguousarray(s.xf[:, :7])).pin_memory()     # position + rotation, 28 B per body
    h_con = torch.empty((w.cfg.maxContacts, 10), dtype=torch.float32).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_loop(upload, sink):
        """One e2e step = upload from pinned memory, the fused step, the contacts in pinned host memory: either fetched
        with axcd_get_contacts after the step (sink False) or delivered by the step itself into the registered
        contact sink (sink True; AxcdStats::numContacts is their count)."""
        w.set_contact_sink(h_con.data_ptr() if sink else None, w.cfg.maxContacts)

        def one():
            upload()                                          # H2D, pinned, inside the timed region
            st2 = w.step()
            nc = st2.numContacts if sink else w.contacts_into(h_con.data_ptr(), w.cfg.maxContacts)   # D2H of the result
            return st2, nc
        for _ in range(2):
            one()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        units = 0
        for _ in range(e2e_steps):
            st2, nc = one()
            units += st2.numPairs + nc
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        units = float(units)
        if world > 1:
            ms = allreduce(torch, dist, [ms], dist.ReduceOp.MAX)[0]
            units = allreduce(torch, dist, [units], dist.ReduceOp.SUM)[0]
        return units / (ms * 1e-3), ms, st2

    # every step uploads whole Transforms (40 B per body) and fetches the contacts with axcd_get_contacts ...
    e2e_full_value, e2e_full_ms, st2 = e2e_loop(lambda: w.set_transforms_ptr(h_xf.data_ptr(), s.n), sink=False)
    # ... or, as a rigid-body step does, position + rotation only (axcd_set_poses, 28 B per body: the scales went to the
    # device with the set_transforms above and do not change from step to step), with the contacts delivered into a
    # registered page-locked buffer while the narrowphase runs (axcd_set_contact_sink: DMA chunks behind the kernel).  This is the `e2e` of the line.
    e2e_value, e2e_ms, st2 = e2e_loop(lambda: w.set_poses_ptr(h_pose.data_ptr(), s.n), sink=True)
    sink_ok = bool(np.array_equal(h_con[:st2.numContacts].numpy().view(np.uint32),
                                  w.contacts().view(np.uint32).reshape(-1, 10)))
    w.set_contact_sink(None, 0)
    h2d = int(s.n) * 28
    d2h = int(st2.numContacts) * 40 + 4 + 128   # contacts + count + stats block

    # ---- roofline: measured FP32 peak, per-stage table, dominant kernel -----------------------------------
    fp32_peak = w.fp32_peak() if rank == 0 else None
    rows, info = stage_table(avg, st_s, hbm_peak, fp32_peak or 1.0, int(generic_pairs_of(s, st_s)))
    for r in rows:
        if r["stage"] == "refit" and refit_dirty > 0:
            r["ms_write_only_flush"] = round(refit_dirty, 4)
            r["frac_of_hbm_peak_write_only_flush"] = round(r["algorithmic_bytes"] / (refit_dirty * 1e-3) / 1e9 / hbm_peak, 4)
    roofline = roofline_of(rows, info, avg, hbm_peak, peak_src, fp32_peak, args.workload,
                           sm_count=torch.cuda.get_device_properties(local).multi_processor_count,
                           sm_mhz=(clocks or {}).get("sm_mhz"))

    single = rank == 0 and world == 1 and args.workload == "headline" and args.flags == 0
    # ---- the SURVEY 8(f) "next" rows on the same scene (rank 0, N=1): manifolds, scene queries, coherence ---
    next_rows = measure_next_rows(w, s, stream, hbm_peak, local) if single and not args.no_next_rows else None
    n_launch = int(st.kernelLaunches)
    w.close()

    # ---- beside the headline (rank 0, N=1): box-box through GJK/EPA, and the hull mix C2 --------------------
    generic = c2 = None
    if single and not args.no_side_workloads:
        side_steps = max(3, min(args.steps, 20))
        generic = measure_side_workload(torch, axcd, s, axcd.FLAG_BOXBOX_GJK_EPA, stream, flush, local, hbm_peak, peak_src,
                                        fp32_peak, "headline", side_steps,
                                        lambda sc, q: q.numPairs * 0.25)   # box-box pairs of a 50/50 mix
        c2 = measure_side_workload(torch, axcd, axcd.config_scene("C2"), 0, stream, flush, local, hbm_peak, peak_src,
                                   fp32_peak, "C2", side_steps, generic_pairs_of)

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle, one full step of the same workload ----------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.flags == 0:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        nthreads = host_threads()
        np_, nc_, _, sec, parts = oracle_step(O, s, nthreads)
        cpu = {"value": (np_ + nc_) / sec, "unit": UNIT, "cores": nthreads, "kind": "port",
               "sample": f"one full step of the same {s.n}-body workload on {nthreads} host threads "
                         f"(refit {parts[0]:.2f}s, grid broadphase {parts[1]:.2f}s, narrowphase {parts[2]:.2f}s)",
               "ms_per_step": 1e3 * sec,
               "pairs_match_gpu": bool(np_ == st.numPairs and nc_ == st.numContacts)}

    # ---- the multi-GPU configurations BASELINE.json names, in the same invocation (N > 1) ------------------
    c3 = c4 = None
    if world > 1 and args.workload == "headline" and not args.no_multi:
        ms_steps = max(3, min(args.steps, 20))
        try:
            c3 = measure_c3(torch, dist, axcd, rank, world, local, stream, flush, ms_steps)
        except Exception as e:   # keep the headline line even if a side configuration fails
            c3 = {"error": repr(e)[:300]}
        try:
            c4 = measure_c4(torch, dist, axcd, rank, world, local, args.scale_c4, ms_steps)
        except Exception as e:
            c4 = {"error": repr(e)[:300]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(args.workload, world, st.numBodies, st.numPairs, st.numContacts, st.numPenetrating),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / e2e_steps,
                    "upload": "axcd_set_poses: position + rotation, 28 B per body from pinned memory (scales resident)",
                    "download": "axcd_set_contact_sink: copy-engine chunks follow the fused narrowphase tile by tile into the pinned host buffer",
                    "sink_matches_device_contacts": sink_ok},
            "e2e_full_transforms": {"value": e2e_full_value, "unit": UNIT, "h2d_bytes_per_step": int(s.n) * 40,
                                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_full_ms / e2e_steps,
                                    "upload": "axcd_set_transforms: whole 40-byte Transforms every step",
                                    "download": "axcd_get_contacts after the step"},
            "gpu_launches": n_launch * args.steps,
            "step_launch": {"graph_launched": graph_launched, "kernels_per_step": n_launch,
                            "ms_per_step_staged_calls": round(staged_total, 4),
                            "note": "value / ms_per_step time the fused step as axcd_step runs it: one CUDA graph "
                                    "launch; the stage table uses the staged calls (direct launches, events between stages)"},
            "clocks": clocks, "roofline": roofline, "stages": rows,
            "fp32_peak_tflops": round(fp32_peak, 2) if fp32_peak else None,
            "target_ms_per_step": 2.0,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if next_rows:
            line["next_rows"] = next_rows
        if generic:
            line["boxbox_through_gjk_epa"] = generic
        if c2:
            line["c2"] = c2
        if c3:
            line["c3"] = c3
        if c4:
            line["c4"] = c4
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="headline", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-next-rows", action="store_true", help="skip the manifold / query / coherence timings")
    ap.add_argument("--no-side-workloads", action="store_true", help="skip the GJK/EPA box-box and C2 lines beside the headline")
    ap.add_argument("--no-multi", action="store_true", help="N>1: skip the c3 / c4 configurations")
    ap.add_argument("--scale", type=float, default=0.125, help="--workload C4 only: fraction of the 16M bodies")
    ap.add_argument("--scale-c4", type=float, default=1.0, help="N>1 default run: fraction of the 16M bodies for the c4 key")
    ap.add_argument("--flags", type=int, default=0, help="AXCD_FLAG_* bits for the context (8 = box-box through GJK/EPA)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
