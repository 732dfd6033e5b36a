# usage: run_vars.sh VARIANT...   (libs prepared as dlux_b200/lib/var_<VARIANT>.so; a trailing T = timing build)
for v in "$@"; do
  cp dlux_b200/lib/var_$v.so dlux_b200/lib/libdlux_b200.so
  echo "=== $v"
  case $v in
    *T) timeout 60 python tools/timing_probe.py 64 2>&1 | grep "MMA\|DRAIN0\|GEN" | tail -8 ;;
    *) timeout 100 python tools/accuracy.py 2>&1 | grep 3xtf32
       timeout 200 python bench.py --steps 300 --warmup 10 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['gemm_ms_per_step'], d['clocks'], d['e2e']['value'])" ;;
  esac
done
cp dlux_b200/lib/var_CUR.so dlux_b200/lib/libdlux_b200.so 2>/dev/null
