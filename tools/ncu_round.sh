#!/bin/bash
# ncu evidence for one round (run under gpurun, ONE GPU): --set full captures of every kernel of the chain at 51 and
# 301 taps, the DRAM traffic of a bench-sized k_pll launch, and the launch list of a short bench.py run.
# Usage: tools/ncu_round.sh <tag>     (writes gpurun_out/<tag>_*.ncu-rep / .csv / .log)
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
for TAPS in 51 301; do
  CMD="python tools/gpu_probe.py --captures 64 --seconds 0.5 --taps $TAPS --kind stereo --reps 2"
  $CMD > $OUT/${TAG}_plain_t$TAPS.log 2>&1 || { echo "plain run failed (taps $TAPS)"; tail -5 $OUT/${TAG}_plain_t$TAPS.log; exit 1; }
  for K in k_rf_demod_win k_bandpass_pair k_audio k_pll; do
    $NCU --set full --import-source on -k regex:$K -s 1 -c 1 -f -o $OUT/${TAG}_${K}_t$TAPS $CMD > $OUT/${TAG}_ncu_${K}_t$TAPS.log 2>&1
    echo "ncu $K taps $TAPS: exit $?"
  done
done
# DRAM traffic of k_pll at the bench's launch size (522 240 steps x 64 captures: 134 MB of trigArg, more than L2)
CMD="python tools/gpu_probe.py --captures 64 --seconds 6 --taps 51 --kind stereo --reps 1"
$CMD > $OUT/${TAG}_plain_big.log 2>&1 &&
$NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum -k regex:k_pll -s 2 -c 1 --csv --log-file $OUT/${TAG}_pll_traffic.csv $CMD > $OUT/${TAG}_ncu_pll_traffic.log 2>&1
echo "ncu pll traffic: exit $?"
# launch list of a short benchmark run
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-extras"
$CMD > $OUT/${TAG}_plain_bench.log 2>&1 &&
$NCU --metrics gpu__time_duration.sum -k regex:^k_ -c 400 --csv --log-file $OUT/${TAG}_launches_bench.csv $CMD > $OUT/${TAG}_ncu_bench.log 2>&1
echo "ncu bench launches: exit $?"
