#!/bin/bash
# ncu evidence for the C3 workload (run under gpurun): one 64M-path wavefront batch (bench.py's batch size) of the 1M-triangle scene.
# The .ncu-rep files stay on the GPU box (tens of MB); their summaries go to gpurun_out/ and from there to profiles/.
set -x
CMD="python tools/bench_render.py --scene mesh -s 32 -b 8 --reps 1 --max-paths 68000000"
$CMD > gpurun_out/plain_c3.log 2>&1 || exit 1
# launch list of the same command (serialised, cold-cache: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c3_final.csv $CMD > gpurun_out/ncu_l.log 2>&1
# full capture of the traversal kernels of the second frame, all 8 depths (the first frame is the occlusion-order learning
# batch: 8 k_extend_rtc launches, its shadow launches are k_shadow_learn)
ncu --set full --clock-control none --import-source on -k regex:"k_shadow_rtc|k_extend_rtc" -s 8 -c 16 -o /tmp/r02_c3_final $CMD > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log
python tools/ncu_summary.py rep /tmp/r02_c3_final.ncu-rep gpurun_out/r02_ncu_c3_final.txt > /dev/null
python tools/make_traffic_json.py /tmp/r02_c3_final.ncu-rep profiles/r02_ncu_c3_final.txt c3 gpurun_out/r02_traffic.json > /dev/null
python tools/ncu_source_regions.py /tmp/r02_c3_final.ncu-rep k_extend_rtc 60 > gpurun_out/r02_ncu_c3_final_extend_lines.txt
python tools/ncu_source_regions.py /tmp/r02_c3_final.ncu-rep k_shadow_rtc 60 > gpurun_out/r02_ncu_c3_final_shadow_lines.txt
# the bandwidth-bound stages of the first two depths of the second frame
ncu --set full --clock-control none -k regex:"k_shade|k_hitinfo|DeviceRadixSortOnesweep" -s 48 -c 12 -o /tmp/r02_c3_final_stages $CMD > gpurun_out/ncu_g.log 2>&1
tail -2 gpurun_out/ncu_g.log
python tools/ncu_summary.py rep /tmp/r02_c3_final_stages.ncu-rep gpurun_out/r02_ncu_c3_final_stages.txt > /dev/null
ls -la /tmp/*.ncu-rep
# build kernels of a 1M-triangle upload: durations only (CAPTURE_BUILD=1)
if [ -n "$CAPTURE_BUILD" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_build_1m.csv -k regex:"k_ploc|k_collapse|k_leaf_boxes|k_pack|k_morton|k_inner|k_select|k_centroid|k_init|DeviceRadixSort|DeviceScan" python tools/bench_trace.py --rays 1000000 --reps 1 > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log
fi
