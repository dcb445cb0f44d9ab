#!/bin/bash
# First GPU call of a round: everything that was built without a GPU (round 1: `-b` barcodes beyond the five goldens of
# tools/bc_smoke.sh, and all of `junctions annotate`) — tests, the annotate side bench with its stage trace, and an ncu
# capture of the new kernels.  Usage on the box:  gpurun --timeout 900 -- tools/gpu_first_call.sh
# Results land in gpurun_out/ (copy what should be judged into profiles/).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
{
  echo "== new GPU test modules"; timeout 600 python -m pytest tests/test_gpu_zz_barcodes.py tests/test_gpu_zzz_annotate.py tests/test_gpu_zzzz_scan_gather.py tests/test_gpu_zzzzz_fuzz.py -q 2>&1 | tail -25
  echo "== cigar_scan A/B: production, launch-bounded configurations 8-10, gather variant 7 (DESIGN 9.1)"
  AB_CASES=5:0:0,5:8:0,5:9:0,5:10:0,7:0:0,7:1:0,7:2:0,7:3:0,7:4:0,5:0:0 timeout 600 python tools/ab_scan.py > $O/ab_scan_r2.json 2> $O/ab_scan_r2.log; tail -12 $O/ab_scan_r2.log
  echo "== annotate side bench (2M junctions, 8 x 30 Mb)"; timeout 900 python tools/bench_annotate.py --steps 3 --warmup 1 2>&1 | tail -3
  d=/tmp/rtjx_bench_annotate/c8_m30
  echo "== annotate stage trace"; RTJX_TRACE=1 regtools_b200/regtools junctions annotate -o /tmp/a.tsv $d/junctions_x*.bed $d/ref.fa $d/ann.gtf 2>&1 | grep "rtjx annotate"
} > $O/first_call.log 2>&1
d=/tmp/rtjx_bench_annotate/c8_m30
timeout 600 ncu --set full --clock-control none --import-source on -k regex:annotate_kernel -c 1 -o $O/annotate_kernel \
  regtools_b200/regtools junctions annotate -o /tmp/a.tsv $d/junctions_x*.bed $d/ref.fa $d/ann.gtf > $O/ncu_annotate.log 2>&1
tools/bamgen gen --out /tmp/sc.bam --config c2 --reads 10000000 --seed 1234 --barcodes 20000 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_barcodes.csv \
  regtools_b200/regtools junctions extract -s XS -b /tmp/sc.bc -o /tmp/sc.bed /tmp/sc.bam > $O/ncu_barcodes.log 2>&1
( time regtools_b200/regtools junctions extract -s XS -b /tmp/sc.bc -o /tmp/sc.bed /tmp/sc.bam ) 2>&1 | grep -v WARNING | tail -5 >> $O/first_call.log
tail -40 $O/first_call.log
