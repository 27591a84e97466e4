#!/bin/bash
# ncu evidence for the C3 workload (run under gpurun): one 32M-path wavefront batch of the 1M-triangle scene.
set -x
CMD="python tools/bench_render.py --scene mesh -s 16 -b 8 --reps 1 --max-paths 34000000"
$CMD > gpurun_out/plain_c3.log 2>&1 || exit 1
# launch list of the same command (serialised, cold-cache: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c3_sorted.csv $CMD > gpurun_out/ncu_l.log 2>&1
# full capture of the second frame's kernels (the first frame is the occlusion-order learning batch)
ncu --set full --clock-control none --import-source on -k regex:"k_shadow_pool|k_shadow_rtc|k_extend_rtc|k_shade|k_hitinfo|k_extend_fallback" -s 40 -c 15 -o gpurun_out/r02_c3_sorted $CMD > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log
# build kernels of a 1M-triangle upload: durations only (CAPTURE_BUILD=1)
if [ -n "$CAPTURE_BUILD" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_build_1m.csv -k regex:"k_ploc|k_collapse|k_leaf_boxes|k_pack|k_morton|k_inner|k_select|k_centroid|k_init|DeviceRadixSort|DeviceScan" python tools/bench_trace.py --rays 1000000 --reps 1 > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log
fi
