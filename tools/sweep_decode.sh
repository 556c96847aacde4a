python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 50 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 600 gpurun_out/bench_1gpu.json
python tools/latency_bench.py > gpurun_out/latency.json 2> gpurun_out/latency.err; tail -c 800 gpurun_out/latency.json
