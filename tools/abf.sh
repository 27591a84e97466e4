#!/bin/bash
# A/B of engine flags / library builds on the 1M-triangle render: tools/abf.sh [lib.so:]flags ...   (B2PT_FLAG_* sum; "default" lib = in-tree)
# Optional env: ABF_ARGS (extra bench_render.py arguments, default "-s 16 -b 8 --max-paths 33554432")
ARGS=${ABF_ARGS:--s 16 -b 8 --max-paths 33554432}
for v in "$@"; do
  lib=${v%%:*}; flags=${v##*:}
  if [ "$lib" = "$v" ] || [ "$lib" = default ]; then unset B2PT_LIB; else export B2PT_LIB=$PWD/$lib; fi
  python tools/bench_render.py --scene mesh $ARGS --reps 2 --flags $flags 2>&1 | tail -1 | python -c "
import sys,json; j=json.loads(sys.stdin.read()); print('$v | mesh %6.1f Msamples/s (ext %6.1f shd %6.1f of %6.1f ms) fallback %d mean %.9f' % (j['msamples_s'], j['extend_ms'], j['shadow_ms'], j['ms'], j['fallback'], j['mean']))"
done
