#!/bin/bash
# Quick GPU check of the -b mode through the CLI against the reference-made goldens (tests/golden/barcodes).
cd "$(dirname "$0")/.." || exit 1
G=tests/golden/barcodes; B=regtools_b200/regtools; T=$(mktemp -d); mkdir -p gpurun_out; L=gpurun_out/bc_smoke.txt; : > $L
run() {  # name bam bed bc args...
  local name=$1 bam=$2 bed=$3 bc=$4; shift 4
  local t0=$(date +%s.%N)
  $B junctions extract "$@" -b $T/o.bc -o $T/o.bed $bam 2> $T/o.err; local rc=$?
  local w=$(grep -c "WARNING: No CB tag" $T/o.err)
  if [ $rc -eq 0 ] && cmp -s $T/o.bed $bed && cmp -s $T/o.bc $bc; then echo "$name OK warnings=$w" | tee -a $L
  else echo "$name FAIL rc=$rc warnings=$w" | tee -a $L; grep -v WARNING $T/o.err | tail -5 | tee -a $L; cmp $T/o.bed $bed 2>&1 | tee -a $L; cmp $T/o.bc $bc 2>&1 | tee -a $L; diff <(head -3 $T/o.bc) <(head -3 $bc) | head -8 | tee -a $L; fi
}
run small $G/bc.bam $G/bc.bed $G/bc.barcodes -s XS
run xs $G/bc2.bam $G/bc2.xs.bed $G/bc2.xs.barcodes -s XS
run xs_a0 $G/bc2.bam $G/bc2.xs_a0.bed $G/bc2.xs_a0.barcodes -s XS -a 0
run rf_m50 $G/bc2.bam $G/bc2.rf_m50.bed $G/bc2.rf_m50.barcodes -s RF -m 50
run fr_region $G/bc2.bam $G/bc2.fr_region.bed $G/bc2.fr_region.barcodes -s FR -r 10:1-60000
