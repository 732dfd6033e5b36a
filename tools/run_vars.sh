# usage: run_vars.sh VARIANT...   (libs prepared as dlux_b200/lib/var_<VARIANT>.so; a trailing T = timing build,
# a trailing S = also the sustained record, a trailing P = run the GPU parity tests with that library: do not
# end a variant NAME in T, S or P)
for v in "$@"; do
  lib=${v%S}; lib=${lib%P}
  cp dlux_b200/lib/var_$lib.so dlux_b200/lib/libdlux_b200.so
  echo "=== $v"
  case $v in
    *T) timeout 60 python tools/timing_probe.py 64 2>&1 | grep "MMA\|DRAIN0\|GEN\|CONV" | tail -10 ;;
    *P) timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ;;
    *S) timeout 200 python bench.py --steps 50 --warmup 5 --sustained 4 --sparse 0 --c4-stars 0 2>/dev/null | tail -1 > /tmp/b.json; python tools/bench_summary.py /tmp/b.json ;;
    *) timeout 90 python bench.py --steps 50 --warmup 5 --sustained 0 --sparse 0 --c4-stars 0 2>/dev/null | tail -1 > /tmp/b.json; python tools/bench_summary.py /tmp/b.json ;;
  esac
done
cp dlux_b200/lib/var_CUR.so dlux_b200/lib/libdlux_b200.so 2>/dev/null
