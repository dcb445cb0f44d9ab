#!/bin/bash
# Developer helper: whole-file identity with the unmodified reference for option sets other than the bench's `-s XS` —
# strand modes RF / FR and non-default anchor / intron bounds — on a generated C3 BAM.   tools/check_modes.sh [reads]
cd "$(dirname "$0")/.." || exit 1
READS=${1:-30000000}
bam=$(python -c "import bench; print(bench.ensure_bam('c3', $READS, 6))" 2>/dev/null | tail -1)
cat $bam > /dev/null
declare -a SETS=("-s RF" "-s FR" "-s XS -a 3 -m 50 -M 100000" "-s RF -a 12 -m 200 -M 20000")
i=0
for s in "${SETS[@]}"; do ( /usr/bin/env bash -c "oracle/_ref/regtools_ref junctions extract $s -o /tmp/modes_ref_$i.bed $bam > /dev/null 2>&1" ) & i=$((i+1)); done
i=0
for s in "${SETS[@]}"; do
  st=$(date +%s.%N); regtools_b200/regtools junctions extract $s -o /tmp/modes_ours_$i.bed $bam > /dev/null 2>&1; rc=$?; en=$(date +%s.%N)
  echo "ours [$s]: rc=$rc $(python -c "print(round($en - $st, 2))") s, $(wc -l < /tmp/modes_ours_$i.bed) lines"; i=$((i+1))
done
wait
i=0
for s in "${SETS[@]}"; do
  if cmp -s /tmp/modes_ours_$i.bed /tmp/modes_ref_$i.bed; then echo "[$s] BED12 identical to the reference ($(sha256sum < /tmp/modes_ref_$i.bed | cut -c1-16), $(wc -l < /tmp/modes_ref_$i.bed) lines)"
  else echo "[$s] DIFFERS"; cmp /tmp/modes_ours_$i.bed /tmp/modes_ref_$i.bed | head -2; fi
  i=$((i+1))
done
