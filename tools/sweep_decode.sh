for cfg in "-DSSE_KSTEPS=1 -DSSE_ROUNDS=16" "-DSSE_KSTEPS=2 -DSSE_ROUNDS=4" "-DSSE_KSTEPS=2 -DSSE_ROUNDS=16" "-DSSE_KSTEPS=2 -DSSE_ROUNDS=8 -DSSE_V3_WARPS=28" "-DSSE_KSTEPS=2 -DSSE_ROUNDS=8 -DSSE_V3_WARPS=20"; do
  SSE_NVCC_DEFS="$cfg" python inference_gateway_b200/build.py --force > /dev/null 2>&1 || echo "build failed $cfg"
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$cfg', d['ms_per_step'])"
done
