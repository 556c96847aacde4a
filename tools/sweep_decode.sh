for cfg in "-DSSE_V3_WARPS=28" "-DSSE_V3_WARPS=30" "-DSSE_V3_WARPS=32" "-DSSE_V3_WARPS=24" "-DSSE_KSTEPS=3" "-DSSE_SKIPW=6" "-DSSE_LEN_SHIFT=4"; do
  SSE_NVCC_DEFS="$cfg" python inference_gateway_b200/build.py --force > /dev/null 2>&1 || echo "build failed $cfg"
  timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$cfg', d['ms_per_step'])"
done
