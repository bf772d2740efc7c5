#!/usr/bin/env python
"""bench.py — the hot path of BASELINE.json on synthetic Nanopore-shaped reads.

  python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, sm_100a)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU LowHash0 (oracle/_ref)

A "step" is one pass of the hot path over one batch: Assembler::findAlignmentCandidatesLowHash0 on the
whole read set (all MinHash iterations). `value` = candidate read pairs emitted per second with the
marker k-mer ids already resident in HBM; `e2e` = the same through the reference-facing C-ABI call with
HOST buffers (7-byte CompressedMarker records, pinned), host->device and device->host copies inside the
timed region. One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# conf/Nanopore-May2022.conf [MinHash] (+ defaults of src/AssemblerOptions.cpp:327-378)
MINHASH_MAY2022 = dict(m=4, hashFraction=0.01, minHashIterationCount=10, alignmentCandidatesPerRead=20.0,
                       log2MinHashBucketCount=0, minBucketSize=5, maxBucketSize=30, minFrequency=5)

WORKLOADS = {
    # BASELINE.json configs[1]: 1M synthetic Nanopore reads (N50 30 kb, ~30x), Nanopore-May2022.conf
    "nanopore-may2022-1M": dict(reads=1_000_000, n50=30_000, coverage=30.0, minhash=MINHASH_MAY2022),
    # Smaller variants for quick runs / N>1 development
    "nanopore-may2022-100k": dict(reads=100_000, n50=30_000, coverage=30.0, minhash=MINHASH_MAY2022),
    "nanopore-may2022-10k": dict(reads=10_000, n50=30_000, coverage=30.0, minhash=MINHASH_MAY2022),
}
CPU_SAMPLE_READS = int(os.environ.get("SHB_CPU_SAMPLE_READS", "40000"))     # bounded sample of the same workload (same coverage, smaller genome)


def synth_params(reads, n50, coverage, seed=1):
    from shasta_b200 import synth
    mean_gap = 13.55
    mean_len = n50 * np.exp(-0.5 * 0.5 ** 2)          # log-normal: mean = N50 * exp(-sigma^2/2)
    span = mean_len / mean_gap
    genome_markers = int(max(reads * span / coverage, 4 * span))
    return synth.SynthParams(reads=reads, k=14, genome_markers=genome_markers, mean_gap=mean_gap, n50_bases=n50,
                             sigma=0.5, min_bases=10_000, drop=0.12, ins=0.05, repeat_period=5000, repeat_len=200,
                             seed=seed)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def run_reference(args, wl, rank, world):
    """The reference's own CPU LowHash0 (oracle/_ref, unmodified TUs; the oracle port if _ref is absent)
    on a bounded sample of the same workload, all host threads."""
    if rank != 0:
        return
    from oracle import bindings as B
    from shasta_b200 import synth
    p = synth_params(CPU_SAMPLE_READS, wl["n50"], wl["coverage"], seed=2)
    d = synth.generate(p)
    cores = os.cpu_count()
    params = B.LowHashParams(**wl["minhash"])
    kind = "reference" if B.have_ref() else "port"

    def once():
        t0 = time.perf_counter()
        if kind == "reference":
            c, _, _, sec = B.ref_lowhash0(d["toc"], d["data"], d["flags"], params, threads=cores)
        else:
            c, _, _ = B.oracle_lowhash0(d["toc"], d["data"], d["flags"], params)
            sec = time.perf_counter() - t0
        return len(c), sec

    for _ in range(args.warmup):
        once()
    total = 0.0
    n = 0
    for _ in range(args.steps):
        n, sec = once()
        total += sec
    value = n * args.steps / total
    M = int(d["toc"][-1])
    line = {
        "impl": "reference", "metric": "lowhash_candidate_read_pairs_per_s", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": args.workload, "minhash": wl["minhash"], "sample": f"{CPU_SAMPLE_READS} reads, {M} markers, same coverage"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores if kind == "reference" else 1, "kind": kind,
                         "sample": f"{CPU_SAMPLE_READS} synthetic reads ({M} markers both strands, {wl['coverage']}x), "
                                   f"{wl['minhash']['minHashIterationCount']} LowHash iterations, {n} candidates",
                         "marker_iterations_per_s": M * wl["minhash"]["minHashIterationCount"] * args.steps / total},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="nanopore-may2022-1M", choices=list(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    import torch
    import torch.distributed as dist
    from shasta_b200 import capi, synth

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic input, generated on the device ------------------------------------------------------
    # Weak scaling: every rank holds its own read set of the workload's size (independent shards).
    p = synth_params(wl["reads"], wl["n50"], wl["coverage"], seed=1 + rank)
    ctx = capi.Context(local_rank)
    t0 = time.perf_counter()
    want_e2e = not args.no_e2e
    dm = capi.synth_generate_device(ctx, p, want_data7=want_e2e)
    M = dm.marker_count
    gen_s = time.perf_counter() - t0
    params = capi.make_lowhash_params(**wl["minhash"])
    ctx.set_markers_device(dm.toc, dm.kmer_ptr, dm.flags, keepalive=dm)

    # ---- kernel-resident timing: inputs already in HBM ---------------------------------------------------
    for _ in range(args.warmup):
        cand, _, _, res = ctx.lowhash0(params, want_stats=True)
    barrier()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    sweep_ms = 0.0
    sweep_launches = 0
    launches = 0
    device_ms = 0.0
    with ClockSampler(local_rank) as clocks:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cand, stats, _, res = ctx.lowhash0(params, want_stats=True)
            sweep_ms += res.sweepMs
            device_ms += res.totalMs
            sweep_launches += res.sweepLaunches
            launches += res.kernelLaunches
        barrier()
        wall = time.perf_counter() - t0
    # Device timing: the library brackets every call with CUDA events on the stream it launches on
    # (res.totalMs, includes the result copies); the wall clock around the K blocking calls is reported beside it.
    wall_host = wall
    wall = 1e-3 * device_ms
    t = torch.tensor([wall], dtype=torch.float64, device="cuda")
    ncand = torch.tensor([float(len(cand))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ncand, op=dist.ReduceOp.SUM)
    wall = float(t.item())
    total_cand = float(ncand.item())
    value = total_cand * args.steps / wall

    # ---- end to end through the reference-facing call with host buffers -----------------------------------
    e2e = None
    if want_e2e:
        host = torch.empty(M * 7, dtype=torch.uint8, pin_memory=True)
        data7 = host.numpy()
        dm.data7_to_host(out=data7)
        e2e_steps = max(2, args.steps)
        ctx.find_alignment_candidates_lowhash0(dm.toc, data7, dm.flags, params)      # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            c2, s2, r2 = ctx.find_alignment_candidates_lowhash0(dm.toc, data7, dm.flags, params)
        barrier()
        e2e_wall = time.perf_counter() - t0
        assert np.array_equal(c2, cand), "host-buffer path and device-resident path disagree"
        t = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_wall = float(t.item())
        e2e = {"value": total_cand * e2e_steps / e2e_wall, "unit": "pairs/s",
               "h2d_bytes_per_step": int(M * 7 + dm.toc.nbytes + dm.flags.nbytes),
               "d2h_bytes_per_step": int(len(c2) * 12 + s2.nbytes), "steps": e2e_steps,
               "ms_per_step": 1e3 * e2e_wall / e2e_steps}

    if rank != 0:
        return

    # ---- roofline of the dominant kernel (hash sweep) --------------------------------------------------
    iters = wl["minhash"]["minHashIterationCount"]
    frac_h = wl["minhash"]["hashFraction"]
    bytes_per_marker_iteration = 4.0 + 16.0 * frac_h
    iters_per_launch = iters * args.steps / max(sweep_launches, 1)
    alg_bytes_per_launch = M * bytes_per_marker_iteration * iters_per_launch
    avg_launch_s = 1e-3 * sweep_ms / max(sweep_launches, 1)
    achieved = alg_bytes_per_launch / avg_launch_s / 1e9
    peak, peak_src = measured_hbm_peak()
    traffic = None
    tp = os.path.join(ROOT, "profiles", "sweep_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "lowhashSweepKernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes_per_launch, "avg_launch_ms": 1e3 * avg_launch_s,
                "iterations_fused_per_launch": iters_per_launch, "sweep_share_of_step": sweep_ms / (1e3 * wall)}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import bindings as B
        ps = synth_params(CPU_SAMPLE_READS, wl["n50"], wl["coverage"], seed=2)
        # Same generator as the numpy one (bit-identical), run on the device to save minutes of host time.
        dms = capi.synth_generate_device(ctx, ps, want_data7=True)
        d = {"toc": dms.toc, "data": dms.data7_to_host(), "flags": dms.flags}
        dms.free()
        cores = os.cpu_count()
        bp = B.LowHashParams(**wl["minhash"])
        if B.have_ref():
            c, _, _, sec = B.ref_lowhash0(d["toc"], d["data"], d["flags"], bp, threads=cores)
            kind = "reference"
        else:
            t0 = time.perf_counter()
            c, _, _ = B.oracle_lowhash0(d["toc"], d["data"], d["flags"], bp)
            sec = time.perf_counter() - t0
            kind, cores = "port", 1
        Ms = int(d["toc"][-1])
        cpu_baseline = {"value": len(c) / sec, "unit": "pairs/s", "cores": cores, "kind": kind,
                        "sample": f"{CPU_SAMPLE_READS} synthetic reads ({Ms} markers both strands, same coverage/config), "
                                  f"{iters} iterations, {len(c)} candidates in {sec:.2f} s",
                        "marker_iterations_per_s": Ms * iters / sec}

    line = {
        "metric": "lowhash_candidate_read_pairs_per_s", "value": value, "unit": "pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "host_wall_ms_per_step": 1e3 * wall_host / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": args.workload, "reads_per_gpu": wl["reads"], "markers_per_gpu": M, "minhash": wl["minhash"],
                   "parallelism": "independent read shards" if world > 1 else "single GPU",
                   "l2": "inputs (k-mer ids %.1f GB) larger than L2" % (4e-9 * M), "generation_s": gen_s},
        "candidates": total_cand, "marker_iterations_per_s": world * M * iters * args.steps / wall,
        "gpu_launches": int(launches), "clocks": clocks.summary(),
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
