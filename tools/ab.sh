#!/bin/bash
# A/B of library builds under gpurun: tools/ab.sh build/libA.so build/libB.so ...   ("default" = the in-tree library)
for lib in "$@"; do
  if [ "$lib" = default ]; then unset B2PT_LIB; else export B2PT_LIB=$PWD/$lib; fi
  m=$(python tools/bench_render.py --scene mesh -s 8 -b 8 --reps 2 2>/dev/null | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('mesh %6.1f (ext %5.1f shd %5.1f ms)' % (j['msamples_s'], j['extend_ms'], j['shadow_ms']))")
  c=$(python tools/bench_render.py --scene cornell --reps 2 2>/dev/null | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('cornell %6.1f (ext %5.1f shd %5.1f ms)' % (j['msamples_s'], j['extend_ms'], j['shadow_ms']))")
  t=$(python tools/bench_trace.py --rays 16000000 --reps 3 2>/dev/null | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('closest %6.1f Mrays/s (%.2f nodes %.2f tris)' % (j['mrays_s_median'], j['nodes_per_ray'], j['tris_per_ray']))")
  a=$(python tools/bench_trace.py --rays 16000000 --reps 3 --any 2>/dev/null | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('any %6.1f' % (j['mrays_s_median']))")
  echo "$lib | $m | $c | $t | $a"
done
