# A/B of decode-kernel build variants (run on the GPU box via gpurun).
for cfg in "-DSSE_ROUNDS=8" "-DSSE_ROUNDS=4" "-DSSE_ROUNDS=16"; do
  SSE_NVCC_DEFS="$cfg" python inference_gateway_b200/build.py --force > /dev/null 2>&1 || echo "build failed $cfg"
  for m in 3; do timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e --mode $m 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$cfg mode $m', d['ms_per_step'])"; done
done
python inference_gateway_b200/build.py --force > /dev/null 2>&1
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
