# usage: run_vars.sh VARIANT[:MODE]...   (libs prepared as dlux_b200/lib/var_<VARIANT>.so)
#   MODE: (none) = bench, S = bench with the sustained record, T = in-kernel timing probe (timing build),
#         P = the GPU parity tests with that library
for arg in "$@"; do
  v=${arg%%:*}; mode=""; case $arg in *:*) mode=${arg##*:};; esac
  cp dlux_b200/lib/var_$v.so dlux_b200/lib/libdlux_b200.so || continue
  echo "=== $arg"
  case $mode in
    T) timeout 60 python tools/timing_probe.py 64 2>&1 | grep "MMA\|DRAIN0\|GEN\|CONV" | tail -10 ;;
    P) timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ;;
    S) timeout 200 python bench.py --steps 50 --warmup 5 --sustained 4 --sparse 0 --c4-stars 0 2>/dev/null | tail -1 > /tmp/b.json; python tools/bench_summary.py /tmp/b.json ;;
    *) timeout 90 python bench.py --steps 50 --warmup 5 --sustained 0 --sparse 0 --c4-stars 0 2>/dev/null | tail -1 > /tmp/b.json; python tools/bench_summary.py /tmp/b.json ;;
  esac
done
cp dlux_b200/lib/var_CUR.so dlux_b200/lib/libdlux_b200.so 2>/dev/null
