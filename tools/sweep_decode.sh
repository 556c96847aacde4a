for cfg in "-DSSE_SKIPW=4" "-DSSE_SKIPW=6" "-DSSE_SKIPW=8"; do
  SSE_NVCC_DEFS="$cfg" python inference_gateway_b200/build.py --force > /dev/null 2>&1 || echo "build failed $cfg"
  timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$cfg', d['ms_per_step'])"
done
python inference_gateway_b200/build.py --force > /dev/null 2>&1
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
