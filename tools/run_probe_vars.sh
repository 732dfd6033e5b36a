for v in "$@"; do
  cp dlux_b200/lib/var_$v.so dlux_b200/lib/libdlux_b200.so
  echo "=== $v"
  timeout 60 python tools/timing_probe.py 1 64 2>&1 | grep "MMA\|DRAIN0" | tail -12
done
