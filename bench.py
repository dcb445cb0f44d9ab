#!/usr/bin/env python
"""bench.py — BAM reads/sec through `junctions extract` (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c3|c2] [--reads R]

Workload (N = 1 and N > 1 alike): BASELINE.json configs[2], the configuration the target is quoted on — a synthetic
100M-read whole-genome BAM (24 contigs, 150 bp paired, 12 % spliced), generated on the box by tools/bamgen (seed 1234,
BGZF level 6) into /dev/shm.  `--config c2` selects configs[1] (10M reads, one chromosome, 101 bp, 8 % spliced) instead.
N > 1 is STRONG scaling on the same file: contigs are sharded over the ranks (one process per GPU), the junction tables
go to rank 0 over NCCL inside the library (rtjx_gather, the path's only exchange), which ranks, sorts and prints.

One "step" = one pass of the hot path over the whole workload.
  value : reads/s with the INPUT RESIDENT IN HBM: the compressed BAM bytes sit in device memory (rtjx_stage_bam) before
          the timed region; a step is BGZF inflate + BAM record split + cigar_scan + junction_merge + finalize (compact,
          first-seen ranking, sort) on the GPU + D2H of the junction table [+ all-gather and merge for N > 1].
  e2e   : reads/s through the public call on the BAM FILE (host buffers: page cache): fresh handle per step, compressed
          bytes staged through pinned memory and copied H2D inside the timed region, BED12 file written and closed —
          wall clock, warm process.  `e2e.cold` = the same through the CLI binary, one process per run, exec to exit.
  roofline : the cigar_scan kernel inside the `value` steps: algorithmic bytes 16*R + 4*C over its CUDA-event duration.
  cpu_baseline : the UNMODIFIED reference (oracle/_ref/regtools_ref, built by oracle/Makefile from /root/reference)
          on one host core (it has no threading), timed on a bounded sample (`-r chr1`) of the same BAM.
  identity : sha256 of the BED12 the timed e2e pass wrote == sha256 of the reference's BED12 on the WHOLE file.

`--impl reference` is the reference arm: nothing of the product is imported or loaded; reads are counted by
oracle/_ref/ref_count (the reference's own htslib).
"""
import argparse
import hashlib
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
REF_COUNT = os.path.join(ROOT, "oracle", "_ref", "ref_count")
ORACLE_BIN = os.path.join(ROOT, "oracle", "_ref", "jx_oracle")
CLI_BIN = os.path.join(ROOT, "regtools_b200", "regtools")
METRIC = "BAM reads/sec through junctions-extract"
SAMPLE_REGION = "chr1"                       # the reference arm's bounded sample: every alignment of the first contig


def bam_path(config, reads, level):
    return os.path.join(SCRATCH, f"rtjx_{config}_{reads}_l{level}.bam")


def ensure_bam(config, reads, level):
    path = bam_path(config, reads, level)
    if not (os.path.exists(path) and os.path.exists(path + ".bai")):
        tmp = path + f".tmp{os.getpid()}.bam"
        subprocess.check_call([BAMGEN, "gen", "--out", tmp, "--config", config, "--reads", str(reads), "--seed", "1234",
                               "--level", str(level)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        os.replace(tmp + ".bai", path + ".bai")
        os.replace(tmp, path)
    return path


def workload_name(config, reads):
    if config == "c2":
        return f"synthetic {reads}-read single-chrom BAM (chr1), 101 bp, ~8% spliced, seed 1234 (BASELINE configs[1])"
    return (f"synthetic {reads}-read whole-genome BAM (24 contigs chr1..chr22,chrX,chrY), 150 bp paired, 12% spliced, seed 1234 "
            "(BASELINE configs[2])")


def sha256_file(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 22), b""):
            h.update(chunk)
    return h.hexdigest()


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


# ---------------------------------------------------------------------------------------------------------------------
# the reference's CPU implementation (oracle/_ref, built from /root/reference by oracle/Makefile; else the oracle port)
# ---------------------------------------------------------------------------------------------------------------------
def ref_cmd(bam, out, region=None, strand="XS"):
    """(argv, kind) of one `junctions extract` run of the reference CPU implementation."""
    r = ["-r", region] if region else []
    if os.path.exists(REF_BIN):
        return [REF_BIN, "junctions", "extract", "-s", strand] + r + ["-o", out, bam], "reference"
    return [ORACLE_BIN, "-s", strand] + r + ["-o", out, bam], "port"


def time_reference(bam, region=None):
    """Wall time of one reference process, exec to exit; returns (seconds, kind, bed_path)."""
    out = os.path.join(SCRATCH, f"rtjx_ref_{os.getpid()}_{'all' if region is None else region.replace(':', '_')}.bed")
    cmd, kind = ref_cmd(bam, out, region)
    t0 = time.perf_counter()
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return time.perf_counter() - t0, kind, out


def count_reads_reference(bam, region):
    """Alignments the reference iterates for `region` — counted by the reference's own htslib (oracle/ref_count.c)."""
    if os.path.exists(REF_COUNT):
        return int(subprocess.check_output([REF_COUNT, bam, region], text=True).split()[0]), "oracle/_ref/ref_count (reference htslib)"
    return None, None


def sidecar_path(bam):
    return bam + ".reference.json"


def whole_file_reference(bam, reads_total):
    """The reference on the WHOLE file, once: its BED12's sha256 and its wall time, cached next to the BAM."""
    sc = sidecar_path(bam)
    if os.path.exists(sc):
        try:
            d = json.load(open(sc))
            if d.get("bam_bytes") == os.path.getsize(bam):
                return d
        except Exception:
            pass
    dt, kind, bed = time_reference(bam, None)
    d = {"sha256": sha256_file(bed), "seconds": dt, "reads": reads_total, "reads_per_s": reads_total / dt, "kind": kind,
         "bed12_bytes": os.path.getsize(bed), "bam_bytes": os.path.getsize(bam),
         "cmd": "regtools junctions extract -s XS -o out.bed <bam>  (whole file, one process, one thread)"}
    os.remove(bed)
    tmp = sc + f".tmp{os.getpid()}"
    json.dump(d, open(tmp, "w"))
    os.replace(tmp, sc)
    return d


def time_reference_all_cores(bam, contigs):
    """SURVEY 8(d)'s courtesy figure: the reference has no threading, so "all host cores" = one reference process per contig
    (`-r chrN`), as many at a time as there are cores, wall = until the last one ends.  NOT output-equivalent (junction names
    restart in every process) — a throughput figure only, labelled as such."""
    cores = min(os.cpu_count() or 1, 64)
    pending = list(contigs)
    running, rcs = [], []
    t0 = time.perf_counter()
    i = 0
    while pending or running:
        while pending and len(running) < cores:
            c = pending.pop(0)
            out = os.path.join(SCRATCH, f"rtjx_refall_{os.getpid()}_{i}.bed")
            i += 1
            cmd, _ = ref_cmd(bam, out, c)
            running.append((subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL), out))
        time.sleep(0.02)
        for p, out in list(running):
            if p.poll() is not None:
                rcs.append(p.returncode)
                running.remove((p, out))
                try:
                    os.remove(out)
                except OSError:
                    pass
    dt = time.perf_counter() - t0
    if any(rcs):
        return None
    return {"seconds": dt, "cores": cores, "processes": len(contigs),
            "note": f"courtesy: one reference process per contig, {cores} at a time; not output-equivalent (per-process junction names)"}


def config_contigs(config):
    return ["chr1"] if config == "c2" else [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY"]


def run_reference_arm(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    n = max(args.gpus, 1)
    bam = ensure_bam(args.config, args.reads, args.level)
    # the whole file once, in the background (one core), while the bounded sample is timed step by step on another
    whole = {}
    th = threading.Thread(target=lambda: whole.update(whole_file_reference(bam, args.reads)), daemon=True)
    th.start()
    sample_reads, counter = count_reads_reference(bam, SAMPLE_REGION)
    times, kind = [], "reference"
    for i in range(args.warmup + args.steps):
        dt, kind, bed = time_reference(bam, SAMPLE_REGION)
        os.remove(bed)
        if i >= args.warmup:
            times.append(dt)
    if sample_reads is None:
        sample_reads = args.reads if args.config == "c2" else None
    ms = 1000.0 * sum(times) / len(times)
    th.join()
    value = (sample_reads / (ms / 1000.0)) if sample_reads else whole["reads_per_s"]
    courtesy = None
    if n == 1 and not args.no_courtesy:
        try:
            courtesy = time_reference_all_cores(bam, config_contigs(args.config))
            if courtesy:
                courtesy["value"] = args.reads / courtesy["seconds"]
                courtesy["unit"] = "reads/s"
        except Exception as e:
            courtesy = {"error": str(e)}
    sample = (f"regtools junctions extract -s XS -r {SAMPLE_REGION}: {sample_reads} reads per step (counted by {counter}); the reference "
              f"is single-threaded (host has {os.cpu_count()} cores)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "reads/s",
        "n_gpus": n, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(args.config, args.reads)},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": 1, "kind": kind, "sample": sample,
                         "whole_file": {k: whole.get(k) for k in ("seconds", "reads", "reads_per_s", "sha256", "bed12_bytes")},
                         "all_cores_courtesy": courtesy},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
def cold_cli_runs(bam, out_bed, n_runs):
    """`regtools junctions extract` as a user runs it: one process per run, wall from exec to exit (CUDA context creation,
    pinned allocations, everything)."""
    times = []
    for _ in range(n_runs):
        t0 = time.perf_counter()
        subprocess.check_call([CLI_BIN, "junctions", "extract", "-s", "XS", "-o", out_bed, bam], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        times.append(time.perf_counter() - t0)
    return times


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=["c3", "c2"])
    ap.add_argument("--reads", type=int, default=0, help="reads in the BAM (default: 100M for c3, 10M for c2)")
    ap.add_argument("--level", type=int, default=6, help="BGZF deflate level of the synthetic BAM")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 5)")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--no-courtesy", action="store_true", help="skip the all-cores courtesy figure")
    ap.add_argument("--no-reference-check", action="store_true", help="skip the whole-file run of the reference when no cached one exists")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if not args.reads:
        args.reads = 100_000_000 if args.config == "c3" else 10_000_000
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import regtools_b200 as rt

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
    if rank == 0:
        ensure_bam(args.config, args.reads, args.level)
    if world > 1:
        dist.barrier()
    bam = bam_path(args.config, args.reads, args.level)
    host_threads = args.threads or max(1, (os.cpu_count() or 1) // world)
    shard = dict(shard_rank=rank, shard_world=world)

    if world > 1:
        # the exchange lives in the library (rtjx_gather: NCCL from HBM to HBM, merge on the root's GPU); torch.distributed only
        # carries the 128-byte NCCL id from rank 0 to the others
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(rt.JunctionsExtractor.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        comm_id = bytes(idt.cpu().numpy().tobytes())

    # ---- value: compressed BAM resident in HBM -> junction table on the host ---------------------------------------
    ex = rt.JunctionsExtractor(bam, ".", 0, "XS", 8, 70, 500000, device=local, profile=True, n_threads=host_threads, **shard)
    if world > 1:
        ex.comm_init(comm_id, rank, world)               # process-wide communicator: later handles reuse it
    ex.stage_bam()

    def step_resident():
        ex.clear()
        ex.identify_junctions_from_BAM()
        if world > 1:
            ex.gather(0)                                 # rank 0 then holds the merged table
        return len(ex.junction_table())

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
    R_rank, C_rank = int(st["reads"]) // args.steps, int(st["cigar_ops"]) // args.steps
    t_ms = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    reads_t = torch.tensor([R_rank], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(reads_t, op=dist.ReduceOp.SUM)
    ms_step = float(t_ms.item()) / args.steps
    reads_all = int(reads_t.item())
    value = reads_all / (ms_step / 1000.0)
    scan_launches = max(int(st["batches"]), 1)
    scan_ms = st["scan_ms"] / scan_launches
    merge_ms = st["merge_ms"] / scan_launches
    resident_stats = {"inflate_kernel_ms_sum_per_step": st["inflate_kernel_ms"] / args.steps,   # summed over launches that run CONCURRENTLY on 4 streams
                       "cigar_scan_ms_per_step": st["scan_ms"] / args.steps,
                      "junction_merge_ms_per_step": st["merge_ms"] / args.steps, "finalize_ms_per_step": st["finalize_ms"] / args.steps,
                      "cigar_scan_launches_per_step": scan_launches / args.steps, "junction_candidates_per_step": int(st["candidates"]),
                      "bgzf_blocks_per_step": int(st["bgzf_blocks"]) // args.steps, "inflated_bytes_per_step": int(st["inflated_bytes"]) // args.steps}
    launches_resident = st["kernel_launches"] / max(args.steps, 1)
    ex.close()

    # ---- end to end through the public call, host buffers ----------------------------------------------------------
    e2e_steps = args.e2e_steps or min(args.steps, 5)
    out_bed = os.path.join(SCRATCH, f"rtjx_bench_{rank}.bed")
    e2e_times, feeder = [], None
    for i in range(1 + e2e_steps):                       # one warm-up (page cache, pinned allocations)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e = rt.JunctionsExtractor(bam, ".", 0, "XS", 8, 70, 500000, device=local, n_threads=host_threads, **shard)
        e.identify_junctions_from_BAM()
        t_run = time.perf_counter()
        if world > 1:
            e.gather(0)
        table = e.junction_table()
        t_fin = time.perf_counter()
        if rank == 0:
            e.output_file_ = out_bed
            e.print_all_junctions()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t1 = time.perf_counter()
        s2 = e.stats()
        e.close()
        if i > 0:
            e2e_times.append(t1 - t0)
            feeder = dict(s2, run_s=t_run - t0, finalize_d2h_s=t_fin - t_run, exchange_merge_bed12_s=t1 - t_fin)
    e2e_t = torch.tensor([sum(e2e_times) / len(e2e_times)], dtype=torch.float64, device=dev)
    h2d_t = torch.tensor([int(feeder["h2d_bytes"]), int(feeder["d2h_bytes"])], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(h2d_t, op=dist.ReduceOp.SUM)
    e2e_value = reads_all / float(e2e_t.item())
    clocks = sampler.stop() if rank == 0 else None       # sampled across both timed regions (resident steps + e2e steps)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- identity: the BED12 of the timed pass against the reference on the whole file ------------------------------
    ours_sha = sha256_file(out_bed)
    identity = {"ours_sha256": ours_sha, "bed12_bytes": os.path.getsize(out_bed), "junction_lines": sum(1 for _ in open(out_bed))}
    if os.path.exists(sidecar_path(bam)) or not args.no_reference_check:
        try:
            whole = whole_file_reference(bam, args.reads)
            identity.update({"reference_sha256": whole["sha256"], "identical_to_reference_on_whole_file": whole["sha256"] == ours_sha,
                             "reference_whole_file_seconds": whole["seconds"], "reference_whole_file_reads_per_s": whole["reads_per_s"],
                             "reference_kind": whole["kind"]})
            if whole["sha256"] != ours_sha:
                sys.stderr.write("bench.py: the BED12 of the timed pass DIFFERS from the reference's on the whole file "
                                 f"({ours_sha[:16]} vs {whole['sha256'][:16]}): the numbers of this line describe a wrong result\n")
        except Exception as ex_:
            identity["reference_error"] = str(ex_)

    # ---- cold: the CLI binary, one process per run (N = 1; the multi-GPU driver is this script) -----------------------
    cold = None
    if world == 1 and os.path.exists(CLI_BIN):
        cold_bed = os.path.join(SCRATCH, "rtjx_bench_cold.bed")
        ct = cold_cli_runs(bam, cold_bed, 2)
        cold = {"value": reads_all / min(ct), "unit": "reads/s", "seconds": ct, "bed12_sha256_equal": sha256_file(cold_bed) == ours_sha,
                "what": "regtools_b200/regtools junctions extract -s XS -o out.bed <bam>: wall from exec to exit, one process per run, page cache warm"}

    # ---- roofline of the dominant kernel of the junction path (cigar_scan) -------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = (16.0 * R_rank + 4.0 * C_rank) / (scan_launches / args.steps)       # per launch
    achieved = alg_bytes / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
    except Exception:
        pass

    # ---- CPU baseline: the reference on a bounded sample of the same BAM (N = 1 only) --------------------------------
    cpu = None
    if n == 1:
        sample_reads, counter = count_reads_reference(bam, SAMPLE_REGION)
        dt, kind, ref_bed = time_reference(bam, SAMPLE_REGION)
        os.remove(ref_bed)
        if sample_reads:
            cpu = {"value": sample_reads / dt, "unit": "reads/s", "cores": 1, "kind": kind,
                   "sample": f"regtools junctions extract -s XS -r {SAMPLE_REGION} on the same BAM: {sample_reads} reads in {dt:.2f} s "
                             f"(counted by {counter}; single-threaded reference; host has {os.cpu_count()} cores)"}
            if "reference_whole_file_reads_per_s" in identity:
                cpu["whole_file_reads_per_s"] = identity["reference_whole_file_reads_per_s"]

    line = {
        "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": n,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(args.config, args.reads), "reads": reads_all,
                   "cigar_ops": C_rank if world == 1 else None, "junctions": int(n_junc), "bgzf_level": args.level,
                   "bam_bytes": os.path.getsize(bam),
                   "l2_policy": "inputs (compressed file, GBs per GPU; SoA batches >= 200 MB) larger than the 126 MB L2; no flush needed",
                   "value_region": "compressed BAM resident in HBM (rtjx_stage_bam) -> BGZF inflate + record split + cigar_scan + junction_merge "
                                   "+ finalize (compact, rank, sort) + D2H of the junction table" + (" (per contig shard) + rtjx_gather: NCCL send/recv of the shard tables to rank 0, ranked and sorted on its GPU" if world > 1 else ""),
                   "host_threads_per_rank": host_threads, "parallelism": f"contig-shard x{world}" if world > 1 else "single GPU"},
        "clocks": clocks,
        "resident": resident_stats,
        "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": int(h2d_t[0].item()), "d2h_bytes_per_step": int(h2d_t[1].item()),
                "ms_per_step": 1000.0 * float(e2e_t.item()), "steps": e2e_steps,
                "region": "fresh handle + rtjx_run (BAM file in the page cache -> pinned staging -> H2D of compressed bytes -> device BGZF inflate "
                          "+ record split + cigar_scan + junction_merge) + finalize + BED12 file written and closed; warm process",
                "feeder": "device" if feeder["host_parse_s"] == 0.0 and feeder["inflated_bytes"] else "host",
                "stages_rank0_s": {"host_staging_memcpy": feeder["host_inflate_s"], "host_wait_for_gpu": feeder["host_wait_s"],
                                   "rtjx_run_total": feeder["run_s"], ("nccl_gather_rank_sort_d2h" if world > 1 else "finalize_and_d2h"): feeder["finalize_d2h_s"],
                                   "bed12_write": feeder["exchange_merge_bed12_s"],
                                   "host_parse": feeder["host_parse_s"]},
                "compressed_bytes_rank0": int(feeder["compressed_bytes"]), "inflated_bytes_rank0": int(feeder["inflated_bytes"]),
                "gpu_launches": int(feeder["kernel_launches"]),
                "cold": cold, "identity": identity},
        "gpu_launches": int(round(launches_resident * args.steps)),
        "roofline": {"bound": "hbm", "kernel": "cigar_scan_small_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None,
                     # ncu's DRAM bytes of one captured launch, scaled to this run's launch size by the algorithmic bytes (same
                     # workload family: traffic is linear in the batch)
                     "traffic": (traffic["cigar_scan_dram_bytes_per_launch"] * alg_bytes / traffic["algorithmic_read_bytes"]
                                 if traffic and traffic.get("algorithmic_read_bytes") else None),
                     "traffic_source": (f"profiles/roofline_traffic.json: {traffic['cigar_scan_dram_bytes_per_launch']:.0f} B read+written by one launch over "
                                        f"{traffic['algorithmic_read_bytes']:.0f} algorithmic bytes ({traffic.get('workload')}; ncu --set full capture, not measured "
                                        "in this run), scaled to this run's launch size" if traffic and traffic.get("algorithmic_read_bytes") else None),
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": scan_ms, "launches_per_step": scan_launches / args.steps,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                     "other_kernels_ms": {"junction_merge": merge_ms, "bgzf_inflate_lanes_sum_over_concurrent_launches_per_step": resident_stats["inflate_kernel_ms_sum_per_step"],
                                          "finalize(compact+rank+sort)_per_step": resident_stats["finalize_ms_per_step"]}},
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
