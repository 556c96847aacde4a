python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 50 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 150 gpurun_out/bench_1gpu.json
python bench.py --steps 50 --warmup 3 --mode 0 --no-cpu-baseline > gpurun_out/bench_1gpu_modeP.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
