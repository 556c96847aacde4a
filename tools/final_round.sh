# Round-end evidence in one GPU call: ncu passes of the shipping code (tools/profile_round.sh), their summaries (written on the
# box, copied to gpurun_out/prof/), then the bench lines, tick, latency and smoke.
# /usr/local/graft/bin/gpurun --timeout 540 -- 'bash tools/final_round.sh r2b'
TAG=${1:-r2b}
mkdir -p gpurun_out/prof
bash tools/profile_round.sh > /dev/null 2>&1
python tools/ncu_summary.py $TAG && cp profiles/${TAG}_* profiles/traffic.json gpurun_out/prof/
python bench.py --steps 50 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 300 gpurun_out/bench_1gpu.json
python tools/tick_bench.py 2>/dev/null | tail -1 > gpurun_out/tick.json; cat gpurun_out/tick.json
python tools/latency_bench.py > gpurun_out/latency.json 2> gpurun_out/latency.err; tail -c 300 gpurun_out/latency.json
python bench.py --steps 50 --warmup 3 --mode 0 --no-cpu-baseline > gpurun_out/bench_1gpu_modeP.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()"
