python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for cfg in "-DSSE_V1_WARPS=10" "-DSSE_V1_WARPS=11" "-DSSE_V1_WARPS=12" "-DSSE_V1_WARPS=12 -DSSE_V3_WARPS=26"; do
  SSE_NVCC_DEFS="$cfg" python inference_gateway_b200/build.py --force > /dev/null 2>&1 || echo "build failed $cfg"
  for m in 3 0; do timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e --mode $m 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$cfg mode $m', d['ms_per_step'])"; done
done
