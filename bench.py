#!/usr/bin/env python
"""bench.py — BAM reads/sec through `junctions extract` (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--reads R]

N = 1 workload (BASELINE.json configs[1]): synthetic 10M-read single-chromosome BAM, 101 bp, ~8 %
spliced, generated on the box by tools/bamgen (seed 1234).  N > 1 (weak scaling): the same shape N
times in one BAM — N contigs of the chr1 length, N x 10M reads — so every GPU does exactly the N = 1
work on its contig shard; junction tables are all-gathered over NCCL (the path's only exchange).

One "step" = one pass of the hot path over the whole workload.
  value : reads/s with the SoA batch already resident in HBM: cigar_scan + junction_merge +
          finalize (compaction, first-seen ranking, sort) + D2H of the junction table
          [+ all-gather for N > 1]; CUDA events on the launching stream, max over ranks.
  e2e   : reads/s through the public call (JunctionsExtractor.identify_junctions_from_BAM +
          print_all_junctions, i.e. rtjx_run + rtjx_write_bed12) on the BAM FILE (host buffers: page
          cache): compressed bytes staged through pinned memory and copied H2D, BGZF inflate + BAM
          record split + CIGAR scan + merge on the GPU, junction table D2H, BED12 file written —
          wall clock, everything inside, a fresh handle per step.
  roofline : cigar_scan kernel, algorithmic bytes 16*R + 4*C over its CUDA-event duration.
  cpu_baseline : the UNMODIFIED reference (oracle/_ref/regtools_ref, built by oracle/Makefile)
          timed on the host cores on a bounded region sample of the same BAM.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCRATCH = os.environ.get("RTJX_SCRATCH", "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp")
BAMGEN = os.path.join(ROOT, "tools", "bamgen")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "regtools_ref")
ORACLE_BIN = os.path.join(ROOT, "oracle", "_ref", "jx_oracle")


def ensure_bam(config, reads, level):
    path = os.path.join(SCRATCH, f"rtjx_{config}_{reads}_l{level}.bam")
    if not (os.path.exists(path) and os.path.exists(path + ".bai")):
        tmp = path + f".tmp{os.getpid()}.bam"
        subprocess.check_call([BAMGEN, "gen", "--out", tmp, "--config", config, "--reads", str(reads), "--seed", "1234",
                               "--level", str(level)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        os.replace(tmp + ".bai", path + ".bai")
        os.replace(tmp, path)
    return path


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).
    NVML from a thread (nvidia-smi -lms buffers its output when piped, which loses every sample on terminate);
    one-shot nvidia-smi calls are the fallback."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index=0, period_s=0.02):
        self.index, self.period, self.rows, self.stop_ev, self.t, self.source = index, period_s, [], threading.Event(), None, None
        self.max_mhz = None

    def _nvml_loop(self, nv, h):
        bits = [getattr(nv, n, 0) for n in ("nvmlClocksEventReasonHwSlowdown", "nvmlClocksEventReasonHwThermalSlowdown",
                                            "nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksEventReasonSwPowerCap")]
        if not any(bits):
            bits = [getattr(nv, n, 0) for n in ("nvmlClocksThrottleReasonHwSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown",
                                                "nvmlClocksThrottleReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwPowerCap")]
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_ev.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = get_reasons(h)
                self.rows.append((int(mhz), [self.NAMES[i] for i in range(4) if bits[i] and (r & bits[i])]))
            except Exception:
                pass
            self.stop_ev.wait(self.period)

    def _smi_loop(self):
        while not self.stop_ev.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split("\n")[0]
                r = [x.strip() for x in o.split(",")]
                self.max_mhz = int(float(r[1]))
                self.rows.append((int(float(r[0])), [self.NAMES[i] for i in range(4) if r[3 + i].lower().startswith("active")]))
            except Exception:
                pass
            self.stop_ev.wait(max(self.period, 0.1))

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.index
            if vis:
                ids = [x for x in vis.split(",") if x.strip()]
                if self.index < len(ids) and ids[self.index].strip().isdigit():
                    phys = int(ids[self.index])
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = int(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.source = "nvml"
            self.t = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
        except Exception:
            self.source = "nvidia-smi"
            self.t = threading.Thread(target=self._smi_loop, daemon=True)
        self.t.start()

    def stop(self):
        if not self.t:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampler not started"], "samples": 0}
        self.stop_ev.set()
        self.t.join(timeout=6)
        sm = sorted(r[0] for r in self.rows)
        reasons = sorted({x for r in self.rows for x in r[1]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm), "source": self.source}


def time_reference(bam, region, threads_note=1, strand="XS"):
    """Wall time of the reference CPU implementation on `region` of `bam`; returns (seconds, kind, bed_path)."""
    out = os.path.join(SCRATCH, f"rtjx_ref_{os.getpid()}.bed")
    if os.path.exists(REF_BIN):
        cmd, kind = [REF_BIN, "junctions", "extract", "-s", strand, "-r", region, "-o", out, bam], "reference"
    else:
        cmd, kind = [ORACLE_BIN, "-s", strand, "-r", region, "-o", out, bam], "port"
    t0 = time.perf_counter()
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return time.perf_counter() - t0, kind, out


def time_reference_all_cores(bam, contig_len, total_reads, contig="chr1"):
    """SURVEY 8(d)'s courtesy figure: the reference has no threading, so "all host cores" = one reference process per window
    of the contig (`-r chr1:a-b`), all started together, wall = the slowest.  NOT output-equivalent (names restart in every
    process, alignments across a window edge are seen twice) — a throughput figure only, labelled as such."""
    cores = min(os.cpu_count() or 1, 64)
    exe = REF_BIN if os.path.exists(REF_BIN) else ORACLE_BIN
    step = -(-contig_len // cores)
    procs = []
    t0 = time.perf_counter()
    for i in range(cores):
        region = f"{contig}:{i * step + 1}-{min((i + 1) * step, contig_len)}"
        out = os.path.join(SCRATCH, f"rtjx_ref_{os.getpid()}_{i}.bed")
        cmd = ([exe, "junctions", "extract"] if exe == REF_BIN else [exe]) + ["-s", "XS", "-r", region, "-o", out, bam]
        procs.append(subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
    rcs = [p.wait() for p in procs]
    dt = time.perf_counter() - t0
    for i in range(cores):
        try:
            os.remove(os.path.join(SCRATCH, f"rtjx_ref_{os.getpid()}_{i}.bed"))
        except OSError:
            pass
    if any(rcs):
        return None
    return {"value": total_reads / dt, "unit": "reads/s", "cores": cores, "seconds": dt,
            "note": f"courtesy: {cores} concurrent reference processes, one per {step}-bp window of {contig}; not output-equivalent "
                    "(per-process junction names, alignments on window edges seen twice)"}


def count_reads(bam, region, device):
    import regtools_b200 as rt
    ex = rt.JunctionsExtractor(bam, region, 0, "XS", 8, 70, 500000, device=device)
    ex.identify_junctions_from_BAM()
    n = ex.stats()["reads"]
    out = os.path.join(SCRATCH, f"rtjx_ours_{os.getpid()}.bed")
    ex.output_file_ = out
    ex.print_all_junctions()
    ex.close()
    return n, out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = max(args.gpus, 1)
    config = "c2" if n == 1 else f"c2x{n}"
    reads = args.reads * n
    bam = ensure_bam(config, reads, args.level)
    # bounded sample: a region holding ~1/5 of a 10M-read workload keeps each step to a few seconds
    region = "chr1:1-50000000"
    import regtools_b200 as rt
    tid, pos, _, _, _ = rt.JunctionsExtractor(bam, region, 0, "XS", 8, 70, 500000, device=-1).load_batch()
    sample_reads = int(len(tid))
    times = []
    kind = "reference"
    for i in range(args.warmup + args.steps):
        dt, kind, _ = time_reference(bam, region)
        if i >= args.warmup:
            times.append(dt)
    ms = 1000.0 * sum(times) / len(times)
    value = sample_reads / (ms / 1000.0)
    courtesy = None
    if n == 1:
        try:
            courtesy = time_reference_all_cores(bam, 248956422, reads)
        except Exception as e:
            courtesy = {"error": str(e)}
    line = {
        "impl": "reference", "metric": "BAM reads/sec through junctions-extract", "value": value, "unit": "reads/s",
        "n_gpus": n, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(config, reads), "sample": f"-r {region} ({sample_reads} reads)"},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": 1, "kind": kind,
                         "sample": f"regtools junctions extract -s XS -r {region}: {sample_reads} reads per step; "
                                   "the reference is single-threaded",
                         "all_cores_courtesy": courtesy},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_name(config, reads):
    if config == "c2":
        return f"synthetic {reads}-read single-chrom BAM (chr1), 101 bp, ~8% spliced, seed 1234 (BASELINE configs[1])"
    if config.startswith("c2x"):
        return (f"synthetic {reads}-read BAM of {config[3:]} contigs, each the configs[1] chromosome (101 bp, ~8% spliced, seed 1234): "
                "one contig shard per GPU, the N=1 work per GPU")
    return f"synthetic {reads}-read whole-genome BAM (24 contigs), 150 bp paired, 12% spliced, seed 1234, contig-sharded"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU")
    ap.add_argument("--level", type=int, default=6, help="BGZF deflate level of the synthetic BAM")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 5)")
    ap.add_argument("--threads", type=int, default=0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import regtools_b200 as rt
    from regtools_b200.distributed import all_gather_tables, merge_tables

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: regtools_b200 has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL writes its log (with NCCL_DEBUG=VERSION/WARN/INFO at least "NCCL version ...") to stdout: keep fd 1 for the ONE
        # JSON line by pointing it at stderr while the communicator is created
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    n = max(args.gpus, world)
    config = "c2" if n == 1 else f"c2x{n}"
    reads_total = args.reads * n
    if rank == 0:
        bam = ensure_bam(config, reads_total, args.level)
    if world > 1:
        dist.barrier()
    bam = os.path.join(SCRATCH, f"rtjx_{config}_{reads_total}_l{args.level}.bam")
    host_threads = args.threads or max(1, (os.cpu_count() or 1) // world)

    # ---- resident batch: this rank's shard as SoA arrays in HBM --------------------------------
    loader = rt.JunctionsExtractor(bam, ".", 0, "XS", 8, 70, 500000, device=local, shard_rank=rank, shard_world=world,
                                   n_threads=host_threads)
    tid, pos, meta, off, cig = loader.load_batch()
    loader.close()
    R, C = int(len(tid)), int(len(cig))
    n_nops = int(np.count_nonzero((cig & 0xF) == 3))
    d = [torch.from_numpy(x.view(np.int32)).to(dev) for x in (tid, pos, meta, off, cig)]
    stream = torch.cuda.current_stream().cuda_stream
    ex = rt.JunctionsExtractor(bam, ".", 0, "XS", 8, 70, 500000, device=local, profile=True)
    from regtools_b200.distributed import _header_contigs
    ex.set_contigs(_header_contigs(bam))

    def step_resident():
        ex.clear()
        ex.scan_batch(*d, first_ordinal=0, n_junction_ops=n_nops, stream=stream)
        ex.finalize(stream)
        t = ex.junction_table()
        if world > 1:
            tabs = all_gather_tables(t, dev)
            return sum(len(x) for x in tabs)
        return len(t)

    for _ in range(args.warmup):
        step_resident()
    ex.reset_stats()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        n_junc = step_resident()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    st = ex.stats()
    t_ms = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    reads_t = torch.tensor([R], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(reads_t, op=dist.ReduceOp.SUM)
    ms_step = float(t_ms.item()) / args.steps
    reads_all = int(reads_t.item())
    value = reads_all / (ms_step / 1000.0)
    scan_ms = st["scan_ms"] / max(st["batches"], 1)
    merge_ms = st["merge_ms"] / max(st["batches"], 1)
    fin_ms = st["finalize_ms"] / max(args.steps, 1)
    launches_resident = st["kernel_launches"] / max(args.steps, 1)
    ex.close()

    # ---- end to end through the public call, host buffers --------------------------------------
    e2e_steps = args.e2e_steps or min(args.steps, 5)
    out_bed = os.path.join(SCRATCH, f"rtjx_bench_{rank}.bed")
    e2e_times, h2d, d2h = [], 0, 0
    for i in range(1 + e2e_steps):                       # one warm-up (page cache, pinned allocs)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e = rt.JunctionsExtractor(bam, ".", 0, "XS", 8, 70, 500000, device=local, shard_rank=rank, shard_world=world,
                                  n_threads=host_threads)
        e.identify_junctions_from_BAM()
        table = e.junction_table()
        if world > 1:
            tabs = all_gather_tables(table, dev)
            if rank == 0:
                m = merge_tables(bam, tabs)
                m.output_file_ = out_bed
                m.print_all_junctions()
                m.close()
        else:
            e.output_file_ = out_bed
            e.print_all_junctions()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dt = time.perf_counter() - t0
        s2 = e.stats()
        e.close()
        if i > 0:
            e2e_times.append(dt)
            h2d, d2h = s2["h2d_bytes"], s2["d2h_bytes"]
            feeder = s2
    e2e_t = torch.tensor([sum(e2e_times) / len(e2e_times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = reads_all / float(e2e_t.item())
    clocks = sampler.stop() if rank == 0 else None       # sampled across both timed regions (resident steps + e2e steps)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (cigar_scan) ------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = 16.0 * R + 4.0 * C
    achieved = alg_bytes / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("cigar_scan_dram_bytes_per_launch")
    except Exception:
        pass

    # ---- CPU baseline: the reference on a bounded sample of the same BAM (N = 1 only) ----------
    cpu = None
    if n == 1:
        region = "chr1:1-50000000"
        sample_reads, ours_bed = count_reads(bam, region, local)
        dt, kind, ref_bed = time_reference(bam, region)
        same = open(ours_bed).read() == open(ref_bed).read()
        cpu = {"value": sample_reads / dt, "unit": "reads/s", "cores": 1, "kind": kind,
               "sample": f"regtools junctions extract -s XS -r {region} on the same BAM: {sample_reads} reads in {dt:.2f} s "
                         f"(single-threaded reference; host has {os.cpu_count()} cores)",
               "bed12_identical_to_ours_on_sample": same}
        try:
            cpu["all_cores_courtesy"] = time_reference_all_cores(bam, 248956422, reads_total)
        except Exception as e:          # the courtesy figure must never cost the bench line
            cpu["all_cores_courtesy"] = {"error": str(e)}

    line = {
        "metric": "BAM reads/sec through junctions-extract", "value": value, "unit": "reads/s", "n_gpus": n,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(config, reads_total), "reads": reads_all, "cigar_ops": C if world == 1 else None,
                   "junction_ops": n_nops if world == 1 else None, "junctions": int(n_junc), "bgzf_level": args.level,
                   "l2_policy": "inputs (>=211 MB per GPU) larger than the 126 MB L2; no flush needed",
                   "value_region": "cigar_scan + junction_merge + finalize (compact, rank, sort) + D2H table"
                                   + (" + NCCL all-gather" if world > 1 else ""),
                   "host_threads_per_rank": host_threads, "parallelism": f"contig-shard x{world}" if world > 1 else "single GPU"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1000.0 * float(e2e_t.item()), "steps": e2e_steps,
                "region": "fresh handle + rtjx_run (BAM file -> pinned staging -> H2D of compressed bytes -> device BGZF inflate "
                          "+ record split + cigar_scan + junction_merge) + finalize + BED12 file write",
                "feeder": "device" if feeder["host_parse_s"] == 0.0 and feeder["inflated_bytes"] else "host",
                "compressed_bytes": int(feeder["compressed_bytes"]), "inflated_bytes": int(feeder["inflated_bytes"]),
                "host_staging_s": feeder["host_inflate_s"], "host_parse_s": feeder["host_parse_s"],
                "gpu_wait_s": feeder["host_wait_s"], "gpu_launches": int(feeder["kernel_launches"])},
        "gpu_launches": int(round(launches_resident * args.steps)),
        "roofline": {"bound": "hbm", "kernel": "cigar_scan_small_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic,
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": scan_ms,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                     "other_kernels_ms": {"junction_merge": merge_ms, "finalize(compact+rank+sort)": fin_ms}},
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
