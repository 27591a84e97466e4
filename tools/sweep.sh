#!/bin/bash
# parameter sweep of the persistent per-lane query kernels (run under gpurun)
for cfg in "2 4 12" "1 4 12" "4 4 12" "8 4 12" "2 1 12" "2 8 12" "2 16 12" "2 4 8" "2 4 16" "4 8 16"; do
  set -- $cfg
  for mode in "" "--any" "--coherent"; do
    B2PT_TPS=$1 B2PT_REFILL=$2 B2PT_BLOCKS_PER_SM=$3 python tools/bench_trace.py --reps 3 $mode 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.readlines()[-1]); print('tps=$1 refill=$2 bps=$3 mode=%-8s %8.1f Mrays/s' % ('$mode' or 'closest', j['mrays_s_median']))"
  done
done
