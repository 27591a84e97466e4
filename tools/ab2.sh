for lib in default build/v_nohoist.so; do
  if [ "$lib" = default ]; then unset B2PT_LIB; else export B2PT_LIB=$PWD/$lib; fi
  python tools/bench_render.py --scene cornell -s 32 --reps 2 2>/dev/null | tail -1 | cut -c1-260
done
