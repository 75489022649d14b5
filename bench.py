#!/usr/bin/env python
"""bench.py — body-steps/s and BH interactions/s of the Barnes-Hut step path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one simulationStep() (nbody_v5_bench.cu:255-283) over the whole body set.
Workloads (BASELINE.json configs): refdisk_1m = configs[1] (1,000,000-body reference disk,
theta 0.5) is the N=1 default; plummer_16m = configs[3]; uniform_16k = configs[0];
plummer_1m = configs[2].  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LAUNCHES_PER_STEP = 18   # graph step: 2 keys + 6 sort + 5 build + 3 centre of mass + 2 force (reset, traversal with the update fused in)
                         # parallel graph branch), force 2, integrate 1 (it also reduces the next step's bounding box);
                         # +2 (reset, bounds) on the first step after an import
FLOP_PER_INTERACTION = 20  # SURVEY §8d (GPU-Gems convention; bench:205-213 op count)

WORKLOADS = {
    "uniform_16k": dict(n=16_384, ic="uniform", desc="configs[0]: 16,384-body uniform cube"),
    "refdisk_1m": dict(n=1_000_000, ic="refdisk", desc="configs[1]: nbody_v5_bench disk (bench:294-308), 1,000,000 bodies"),
    "refdisk_500k": dict(n=500_000, ic="refdisk", desc="reference code default N=500,000 (bench:31)"),
    "plummer_1m": dict(n=1_000_000, ic="plummer", desc="configs[2]: 1,000,000-body Plummer sphere a=200 cut 10a"),
    "plummer_16m": dict(n=16_000_000, ic="plummer", desc="configs[3]: 16,000,000-body Plummer sphere a=200 cut 10a"),
    "twodisk_16m": dict(n=16_000_000, ic="twodisk", desc="two-galaxy collision, 16,000,000 bodies (small version of configs[4])"),
    "twodisk_256m": dict(n=256_000_000, ic="twodisk", desc="configs[4]: 256,000,000-body two-galaxy collision"),
}


def workload_config(args, w):
    """`config` of the JSON line: the WORKLOAD only, identical for both arms (engine knobs live in `engine`)."""
    return {"workload": args.workload, "desc": w["desc"], "n_bodies": w["n"], "theta": 0.5, "G": 0.5, "dt": 0.02,
            "softening": 50.0, "max_speed": 500.0}


def reference_ic(n):
    """The reference disk from the reference's OWN generator (its main() run up to the uploads,
    oracle/ref_wrap.cu:ref_ic) — the reference arm makes its input without touching the engine's libraries."""
    import ctypes as C

    import numpy as np

    path = os.path.join(ROOT, "oracle", "_ref", "libref_step.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    a = [np.zeros(n, np.float32) for _ in range(7)]
    fd = os.dup(1)                      # main() prints its banner: keep it off our stdout
    os.dup2(2, 1)
    try:
        rc = L.ref_ic(n, *[x.ctypes.data_as(C.c_void_p) for x in a])
    finally:
        os.dup2(fd, 1)
        os.close(fd)
    return a if rc == 0 else None


def make_ic(bh, w, reference_generator=False):
    if w["ic"] == "refdisk":
        if reference_generator:
            a = reference_ic(w["n"])
            if a is not None:
                return a
        return bh.ic_refdisk(w["n"], 42)
    if w["ic"] == "uniform":
        return bh.ic_uniform_cube(w["n"], 42, 1000.0)
    if w["ic"] == "twodisk":
        return bh.ic_two_disks(w["n"], 42, 4000.0, 20.0, 8.0)
    return bh.ic_plummer(w["n"], 42, 200.0, 10.0, 4.5, 0.5)


def force_roofline(bh, device, interactions_per_gpu, force_ms):
    """roofline object of the dominant kernel for the multi-GPU lines: per-GPU FP32 rate of the traversal
    (20 flop per accepted interaction) on the slowest rank against the FMA issue-rate probe."""
    try:
        peak = bh.probe_fp32_tflops(device)
        achieved = FLOP_PER_INTERACTION * interactions_per_gpu / (force_ms * 1e-3) / 1e12
        return {"bound": "fp32", "kernel": "force_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak if peak else None, "traffic": None, "per": "GPU (mean interactions / slowest rank's force phase)",
                "peak_source": "bh_probe_fp32_tflops (FMA issue-rate probe run in this process)",
                "flop_per_interaction": FLOP_PER_INTERACTION}
    except Exception as e:   # the line must not be lost over a diagnostic
        return {"error": repr(e)}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_baseline(w, budget_steps=None, reference_generator=False):
    """Oracle-L (OpenMP transliteration of the reference kernels, contract baseline (a)) on the host cores."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import nbody_barnes_hut_cuda_b200 as bh   # only its host-side IC library (libbh_ic.so) is loaded here
    import oracle_lib as O

    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    O.lib().orc_set_num_threads(cores)        # torchrun exports OMP_NUM_THREADS=1: ask for the cores we report
    n = min(w["n"], 1_000_000)
    soa = make_ic(bh, dict(w, n=n), reference_generator)
    steps = budget_steps or (10 if n <= 100_000 else 4)
    O.reference_step(soa, 1, fixed=0)  # warm the allocator / page cache
    t0 = time.perf_counter()
    r = O.reference_step(soa, steps, fixed=0)
    dt = time.perf_counter() - t0
    return {"value": n * steps / dt, "unit": "body-steps/s", "cores": O.num_threads(), "omp_num_threads_set": cores, "kind": "port",
            "sample": f"Oracle-L (literal OpenMP transliteration, 1.00 interactions/body) {steps} steps at N={n} of the {w['ic']} input",
            "phase_ms_per_step": {k: round(v / steps, 3) for k, v in zip(("keys", "sort", "insert", "com", "force", "integrate"), r["phase_ms"])}}


def run_reference_arm(args, w):
    """The reference's own kernels + simulationStep() (sm_100 recompile, oracle/_ref) when a GPU and the
    prebuilt library are present — BASELINE.md B1 / north_star baseline (b); otherwise the OpenMP port."""
    import ctypes as C

    import numpy as np

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    import nbody_barnes_hut_cuda_b200 as bh

    cpu = cpu_baseline(w, reference_generator=True)
    path = os.path.join(ROOT, "oracle", "_ref", "libref_step.so")
    have_gpu = False
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:
        pass
    line = {"impl": "reference", "metric": "body-steps/s", "unit": "body-steps/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, w)}
    if have_gpu and os.path.exists(path):
        L = C.CDLL(path)
        soa = make_ic(bh, w, reference_generator=True)
        line["input"] = ("the reference's own main() run up to its uploads (oracle/ref_wrap.cu:ref_ic)" if w["ic"] == "refdisk"
                         else "libbh_ic.so (host-only IC generators; the reference has no such input of its own)")
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = L.ref_init(w["n"], *[p(np.ascontiguousarray(a, np.float32)) for a in soa])
        assert rc == 0, f"ref_init {rc}"
        ms = C.c_float()
        clocks = ClockSampler()
        # the reference needs ~50 ms/step at 1M and seconds per step at 16M (one-thread bounding box,
        # N/1024 insert launches): bound the sample so the arm ends within minutes
        steps = args.steps if w["n"] <= 1_000_000 else min(args.steps, 5)
        warm = args.warmup if w["n"] <= 1_000_000 else 3
        assert L.ref_step(warm, C.byref(ms)) == 0
        clocks.start()
        assert L.ref_step(steps, C.byref(ms)) == 0
        ck = clocks.stop()
        L.ref_free()
        value = w["n"] * steps / (ms.value * 1e-3)
        line["steps"], line["warmup"] = steps, warm
        line.update({"value": value, "ms_per_step": ms.value / steps, "interactions_per_body": 1.0,
                     "interactions_per_s": value, "clocks": ck, "gpu_launches": None,
                     "reference_kind": "UNMODIFIED nbody_v5_bench.cu kernels + simulationStep() compiled -arch=sm_100 "
                                       "(oracle/ref_wrap.cu includes the source in place), run on this B200; "
                                       "the reference is CUDA-only, its OpenMP transliteration is in cpu_baseline",
                     "cpu_baseline": cpu,
                     "e2e": {"value": value, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    else:
        line.update({"value": cpu["value"], "ms_per_step": 1e3 * min(w["n"], 1_000_000) / cpu["value"],
                     "interactions_per_body": 1.0, "cpu_baseline": cpu,
                     "reference_kind": "OpenMP transliteration (no GPU or oracle/_ref missing)",
                     "e2e": {"value": cpu["value"], "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    return line


def run_ours(args, w):
    import numpy as np
    import torch

    import nbody_barnes_hut_cuda_b200 as bh

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    if world > 1:
        mode = args.mode or ("let" if w["n"] > 100_000_000 else "sliced")   # north_star: replicate up to ~100M bodies
        if mode == "let":
            from nbody_barnes_hut_cuda_b200.let import run_let_bench      # locally-essential-tree exchange

            return run_let_bench(args, w, bh, dist, rank, world, local)
        from nbody_barnes_hut_cuda_b200.sliced import run_sliced_bench  # multi-GPU Morton slices

        return run_sliced_bench(args, w, bh, dist, rank, world, local)

    n = w["n"]
    soa = make_ic(bh, w)
    stream = torch.cuda.current_stream().cuda_stream
    key_bits = args.key_bits or 30
    eng = bh.BHEngine(n, device=local, key_bits=key_bits)
    eng.load_soa(*soa)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()

    # ---- warm-up
    eng.simulation_step(args.warmup, stream)
    barrier()
    eng.check_device_error()

    # ---- timed: K steps, each bracketed by events, L2 flushed (untimed) between steps
    clocks = ClockSampler(local)
    clocks.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    inter = 0
    barrier()
    t_wall0 = time.perf_counter()
    for a, b in evs:
        flush.fill_(1)
        a.record()
        eng.simulation_step(1, stream)
        b.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = sum(step_ms)
    inter = eng.stat(bh.STAT.INTERACTIONS_CELL) + eng.stat(bh.STAT.INTERACTIONS_BODY)

    # ---- same loop back to back (no flush): the natural simulation loop, for information
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    a.record()
    eng.simulation_step(args.steps, stream)
    b.record()
    barrier()
    warm_ms = a.elapsed_time(b)
    ck = clocks.stop()
    eng.check_device_error()
    cells = eng.stat(bh.STAT.CELLS)
    max_stack = eng.stat(bh.STAT.MAX_STACK)
    eng.close()

    # ---- per-phase breakdown (event between phases, direct launches) + roofline of the dominant kernel
    engt = bh.BHEngine(n, device=local, flags=2, key_bits=key_bits)
    engt.load_soa(*soa)
    engt.simulation_step(args.warmup, stream)
    psteps = max(3, min(args.steps, 20))
    engt.simulation_step(psteps, stream)
    phases = {k: v / psteps for k, v in engt.phase_ms().items()}
    inter_t = engt.stat(bh.STAT.INTERACTIONS_CELL) + engt.stat(bh.STAT.INTERACTIONS_BODY)   # of the steps just timed
    engt.close()
    fp32_peak = bh.probe_fp32_tflops(local)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    force_tflops = FLOP_PER_INTERACTION * inter_t / (phases["force"] * 1e-3) / 1e12
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "force_traffic.json"))).get(args.workload)
    except Exception:
        pass
    bytes_per_body = {"keys": 32.0, "sort": 68.0 + 64.0, "update": 60.0}   # SURVEY §8d algorithmic bytes
    hbm_frac = {k: (bytes_per_body[k] * n / (phases[k] * 1e-3) / 1e9) / hbm_peak for k in bytes_per_body}

    # ---- e2e: host SoA in (pinned) -> 1 step -> host SoA out, every copy inside the timed region
    enge = bh.BHEngine(n, device=local, key_bits=key_bits)
    pinned = [torch.from_numpy(x.copy()).pin_memory() for x in soa]
    harr = [t.numpy() for t in pinned]
    for _ in range(max(3, args.warmup)):
        enge.step_host(*harr, nsteps=1)
    esteps = max(3, min(args.steps, 20))
    barrier()
    t0 = time.perf_counter()
    for _ in range(esteps):
        enge.step_host(*harr, nsteps=1)
    barrier()
    e2e_s = (time.perf_counter() - t0) / esteps
    enge.close()

    value = n * args.steps / (total_ms * 1e-3)
    line = {
        "metric": "body-steps/s", "value": value, "unit": "body-steps/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, w),
        "engine": {"group": 32, "key_bits": key_bits, "launches_per_step": LAUNCHES_PER_STEP,
                   "l2": "flushed between timed steps (256 MiB fill, untimed); value_l2_warm is the back-to-back loop"},
        "interactions_per_body": inter / n, "interactions_per_s": inter * args.steps / (total_ms * 1e-3),
        "value_l2_warm": n * args.steps / (warm_ms * 1e-3), "ms_per_step_l2_warm": warm_ms / args.steps,
        "wall_s_timed_loop": t_wall, "cells": cells, "max_stack": max_stack,
        "phase_ms": {k: round(v, 4) for k, v in phases.items()},
        "roofline": {"bound": "fp32", "kernel": "force_kernel", "achieved": force_tflops, "peak": fp32_peak,
                     "unit": "TFLOP/s", "frac": force_tflops / fp32_peak if fp32_peak else None, "traffic": traffic,
                     "traffic_source": "STATIC: dram bytes read+write of one force_kernel launch from the committed ncu --set full "
                                       "capture (profiles/force_traffic.json names the report); not measured in this run",
                     "peak_source": "bh_probe_fp32_tflops (FMA issue-rate probe run in this process)",
                     "flop_per_interaction": FLOP_PER_INTERACTION,
                     "hbm_frac_streaming_phases": {k: round(v, 4) for k, v in hbm_frac.items()},
                     "hbm_peak_gbs": hbm_peak, "hbm_peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
        "e2e": {"value": n / e2e_s, "unit": "body-steps/s", "h2d_bytes_per_step": 28 * n, "d2h_bytes_per_step": 24 * n,
                "ms_per_step": e2e_s * 1e3, "api": "bh_step_host (pinned host SoA in, 1 step, host SoA out)"},
        "gpu_launches": LAUNCHES_PER_STEP * args.steps, "clocks": ck,
    }
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(w)
    if args.workload == "refdisk_1m" and not args.no_scale_ref:
        # the N>1 runs use the 16M-body Plummer sphere (BASELINE.json configs[3]); measure it on this one GPU
        # too so the 1/2/4/8-GPU series can be read off the same workload
        w2 = WORKLOADS["plummer_16m"]
        soa2 = make_ic(bh, w2)
        eng2 = bh.BHEngine(w2["n"], device=local)
        eng2.load_soa(*soa2)
        eng2.simulation_step(3, stream)
        barrier()
        a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a2.record()
        eng2.simulation_step(10, stream)
        b2.record()
        barrier()
        eng2.check_device_error()
        ms2 = a2.elapsed_time(b2) / 10
        inter2 = eng2.stat(bh.STAT.INTERACTIONS_CELL) + eng2.stat(bh.STAT.INTERACTIONS_BODY)
        eng2.close()
        line["scale_ref_1gpu"] = {"workload": "plummer_16m", "n_bodies": w2["n"], "ms_per_step": ms2,
                                  "value": w2["n"] / (ms2 * 1e-3), "unit": "body-steps/s",
                                  "interactions_per_body": inter2 / w2["n"],
                                  "note": "same workload as the default --gpus 2/4/8 runs (strong scaling)"}
    return line


def main():
    # keep stdout for the ONE JSON line: libraries (NCCL's version banner, torchrun notices) get stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS))
    ap.add_argument("--mode", default=None, choices=["sliced", "let"],
                    help="multi-GPU scheme: replicated tree + Morton slices, or locally-essential-tree exchange")
    ap.add_argument("--key-bits", type=int, default=None, choices=[30, 60],
                    help="Morton key width (bh_params.key_bits); default 30 = the reference key, 60 above 100M bodies")
    ap.add_argument("--let-interval", type=int, default=None, help="LET mode: steps between cube/election/migration rounds")
    ap.add_argument("--let-no-rebalance", action="store_true", help="LET mode: equal-count key ranges instead of equal work")
    ap.add_argument("--no-e2e", action="store_true", help="LET mode: skip the host-buffer leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scale-ref", action="store_true", help="skip the 16M-body single-GPU reference point")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.workload is None:
        args.workload = "refdisk_1m" if args.gpus == 1 else "plummer_16m"
    w = WORKLOADS[args.workload]
    line = run_reference_arm(args, w) if args.impl == "reference" else run_ours(args, w)
    if line is not None:
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()


if __name__ == "__main__":
    main()
