#!/bin/bash
# Developer helper: one `-r <contig>` run of the CLI on a generated C3 BAM, the unmodified reference beside it, outputs compared.
#   tools/bench_region.sh [reads] [region]
cd "$(dirname "$0")/.." || exit 1
READS=${1:-30000000}; REG=${2:-chr1}
bam=$(python -c "import bench; print(bench.ensure_bam('c3', $READS, 6))" 2>/dev/null | tail -1)
cat $bam > /dev/null
t() { local s=$(date +%s.%N); "$@" > /dev/null 2>&1; local rc=$?; local e=$(date +%s.%N); python -c "print(round($e - $s, 3), 's rc=$rc')"; }
echo "reads in $REG: $(oracle/_ref/ref_count $bam $REG 2>/dev/null | cut -f1)"
echo -n "ours (auto: device feeder for large regions), run 1: "; t regtools_b200/regtools junctions extract -s XS -r $REG -o /tmp/reg_ours.bed $bam
echo -n "ours, run 2: "; t regtools_b200/regtools junctions extract -s XS -r $REG -o /tmp/reg_ours.bed $bam
echo "trace of a third run:"; RTJX_TRACE=1 regtools_b200/regtools junctions extract -s XS -r $REG -o /tmp/reg_ours.bed $bam 2>&1 | grep "^\[rtjx" | cut -c1-400
echo -n "ours with RTJX_REGION_DEVICE_MB=1000000 (host reader, the round-1 path): "; RTJX_REGION_DEVICE_MB=1000000 t regtools_b200/regtools junctions extract -s XS -r $REG -o /tmp/reg_host.bed $bam
if [ -x oracle/_ref/regtools_ref ]; then
  echo -n "reference: "; t oracle/_ref/regtools_ref junctions extract -s XS -r $REG -o /tmp/reg_ref.bed $bam
  cmp /tmp/reg_ours.bed /tmp/reg_ref.bed && echo "BED12 identical to the reference ($(wc -l < /tmp/reg_ref.bed) lines)"
fi
cmp /tmp/reg_ours.bed /tmp/reg_host.bed && echo "device-feeder and host-reader BED12 identical"
