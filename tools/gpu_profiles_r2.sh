#!/bin/bash
# Round-2 profiling pass (one GPU): everything the numbers in DESIGN.md / profiles/ come from.
#   gpurun --timeout 2700 -- tools/gpu_profiles_r2.sh [bench|ncu|side|sanitizer ...]   (default: bench ncu side)
# Results land in gpurun_out/ (<= 64 MiB come back: ncu reports are reduced to CSV on the box, only the two headline
# kernels' reports are kept); what should be judged is copied into profiles/ afterwards.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
WHAT="${*:-bench ncu side}"
reduce() {   # reduce() report: raw metrics + per-SASS-instruction table as CSV, then drop the (large) report
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1.raw.csv 2>/dev/null
  ncu -i $O/$1.ncu-rep --page source --csv > $O/$1.source.csv 2>/dev/null
  rm -f $O/$1.ncu-rep
}
{
  if [[ $WHAT == *bench* ]]; then
    echo "== bench, reference arm then ours (C3, 100M reads)"
    timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_c3_ref.json 2> $O/r2_bench_c3_ref.err; tail -c 300 $O/r2_bench_c3_ref.json; echo
    timeout 900 python bench.py > $O/r2_bench_c3.json 2> $O/r2_bench_c3.err; cut -c1-300 $O/r2_bench_c3.json; echo
    echo "== ncu launch list of the bench command"
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-courtesy > $O/r2_bench_under_ncu.log 2>&1
    wc -l $O/r2_launches.csv
  fi
  if [[ $WHAT == *cold* ]]; then
    echo "== cold CLI runs with the engine's trace (exec to exit, page cache warm)"
    bam=$(python -c "import bench; print(bench.ensure_bam('c3', 100000000, 6))" 2>/dev/null | tail -1)
    cat $bam > /dev/null
    for i in 1 2 3; do s=$(date +%s.%N); RTJX_TRACE=1 regtools_b200/regtools junctions extract -s XS -o /tmp/cold.bed $bam 2>&1 | cut -c1-400; e=$(date +%s.%N); echo "cold run $i: $(python -c "print(round($e - $s, 3))") s wall"; done
  fi
  if [[ $WHAT == *ncu* ]]; then
    echo "== ncu --set full: cigar_scan + junction_merge on a 30M-read C3 batch, the lane decoder on a 13M-read C3 file in one launch (57k blocks, 12 warps per SM), feed kernels on the C2 file"
    AB_CASES=5:0:0 timeout 400 ncu --set full --clock-control none -k regex:cigar_scan -s 6 -c 1 -o $O/r2_cigar_scan python tools/ab_scan.py 30000000 6 c3 > $O/ncu_scan.log 2>&1
    reduce r2_cigar_scan
    AB_CASES=5:0:0 timeout 400 ncu --set full --clock-control none -k regex:junction_merge -s 6 -c 1 -o $O/r2_junction_merge python tools/ab_scan.py 30000000 6 c3 > $O/ncu_merge.log 2>&1; reduce r2_junction_merge
    timeout 400 python tools/prof_inflate.py 13000000 0 c3 > $O/r2_inflate_standalone.txt 2>&1; tail -2 $O/r2_inflate_standalone.txt
    RTJX_INFLATE_VARIANT=3 timeout 400 ncu --set full --clock-control none -k regex:bgzf_inflate_lanes -c 1 -o $O/r2_inflate_lanes python tools/prof_inflate.py 13000000 0 c3 > $O/ncu_inflate.log 2>&1
    reduce r2_inflate_lanes
    timeout 400 ncu --set full --clock-control none -k "regex:bgzf_match_resolve|block_seeds|record_walk|record_extract|record_gather" -s 10 -c 5 -o $O/r2_feed_kernels python tools/prof_e2e.py 10000000 0 1 > $O/ncu_feed.log 2>&1; reduce r2_feed_kernels
  fi
  if [[ $WHAT == *side* ]]; then
    echo "== C4 shape: 50k variant windows on the 100M-read BAM"
    timeout 900 python tools/bench_regions.py 100000000 50000 c3 > $O/r2_regions_c4.json 2> $O/r2_regions_c4.err; cat $O/r2_regions_c4.json; tail -3 $O/r2_regions_c4.err
    echo "== junctions annotate side bench + ncu"
    timeout 900 python tools/bench_annotate.py --steps 3 --warmup 1 > $O/r2_annotate.json 2> $O/r2_annotate.err; cat $O/r2_annotate.json; tail -2 $O/r2_annotate.err
    d=/tmp/rtjx_bench_annotate/c8_m30
    timeout 400 ncu --set full --clock-control none -k regex:annotate_kernel -c 1 -o $O/r2_annotate_kernel regtools_b200/regtools junctions annotate -o /tmp/a.tsv $d/junctions_x*.bed $d/ref.fa $d/ann.gtf > $O/ncu_annotate.log 2>&1; reduce r2_annotate_kernel
  fi
  if [[ $WHAT == *side* ]]; then
    echo "== -b single-cell mode at scale (10M reads, 5000 barcodes) through the CLI, reference on a region beside it"
    timeout 600 python tools/bench_barcodes.py 10000000 5000 > $O/r2_barcodes.json 2> $O/r2_barcodes.err; cat $O/r2_barcodes.json; tail -2 $O/r2_barcodes.err
  fi
  if [[ $WHAT == *tests* ]]; then
    echo "== full GPU suite + smoke()"
    timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee $O/r2_pytest_gpu_final.log
    timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a $O/r2_pytest_gpu_final.log
  fi
  if [[ $WHAT == *sanitizer* ]]; then
    echo "== compute-sanitizer (memcheck, racecheck) over the kernel-level parity tests at small sizes"
    timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scan_pipe.py tests/test_gpu_inflate.py -q -x -k "random_batches or ring_configs or fixture_bams or stored_fixed or hot" > $O/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $O/r2_sanitizer_memcheck.log
    timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scan_pipe.py tests/test_gpu_inflate.py -q -x -k "random_batches or ring_configs or fixture_bams or hot" > $O/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 $O/r2_sanitizer_racecheck.log
    timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_device_feed.py tests/test_gpu_regions.py -q -x > $O/r2_sanitizer_memcheck_feed.log 2>&1; echo "memcheck feed rc=$?"; tail -4 $O/r2_sanitizer_memcheck_feed.log
  fi
  rm -f $O/*.ncu-rep; ls -laS $O | head -12; du -sh $O
} > $O/r2_profiles.log 2>&1
# never come home empty-handed: drop the largest files until the directory fits the 64 MiB return limit
while [ "$(du -sm $O | cut -f1)" -gt 55 ]; do f=$(ls -S $O | head -1); echo "dropping $O/$f ($(du -sh $O/$f | cut -f1))" >> $O/r2_profiles.log; rm -f "$O/$f"; done
tail -40 $O/r2_profiles.log | cut -c1-500
