# Round-end validation on the GPU box: /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_validate.sh'
# (numbers printed under ncu by tools/profile_round.sh are never bench values)
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 50 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 150 gpurun_out/bench_1gpu.json
python bench.py --steps 50 --warmup 3 --mode 0 --no-cpu-baseline > gpurun_out/bench_1gpu_modeP.json 2>/dev/null
python tools/latency_bench.py > gpurun_out/latency.json 2> gpurun_out/latency.err
python tools/tick_bench.py 2>/dev/null | tail -1 > gpurun_out/tick.json
python -c "import __graft_entry__ as g; g.smoke()"
# build-time knobs can be A/B-tested with SSE_NVCC_DEFS="-DSSE_V3_WARPS=28 ..." python inference_gateway_b200/build.py --force
