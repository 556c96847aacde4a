python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 50 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 200 gpurun_out/bench_1gpu.json
python bench.py --steps 50 --warmup 3 --mode 0 --no-cpu-baseline > gpurun_out/bench_1gpu_modeP.json 2>/dev/null
python tools/latency_bench.py > gpurun_out/latency.json 2> gpurun_out/latency.err
python tools/tick_bench.py 2>/dev/null | tail -1 > gpurun_out/tick.json; cat gpurun_out/tick.json
python -c "import __graft_entry__ as g; g.smoke()"
bash tools/profile_round.sh > /dev/null 2>&1
