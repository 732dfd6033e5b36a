for mb in 0 1600 800 420 300 220 120; do
  echo "=== CHUNK_MB=$mb"
  DLUX_B200_CHUNK_MB=$mb timeout 200 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['gemm_ms_per_step'], d['roofline']['gemm_launches_per_step'], d['clocks']['sm_mhz'], d['clocks']['power_w'])"
done
