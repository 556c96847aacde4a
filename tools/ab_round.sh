# One GPU call: parity of the default build, A/B of the build-time variants (tools/build_variants.sh) on the C4 bench, launch list
# of the default build, ncu captures of the decode kernels of the named variants ("default" = the in-tree library).
# /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/ab_round.sh [variants to profile]'
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -x -q -m gpu ) > gpurun_out/gputests.log 2>&1; tail -4 gpurun_out/gputests.log | head -2
Q="--steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-strong --segments 0"
python bench.py $Q > gpurun_out/var_default.json 2> gpurun_out/var_default.err
for v in variants/*.so; do n=$(basename $v .so); SSE_LIB=$PWD/$v python bench.py $Q > gpurun_out/var_$n.json 2> gpurun_out/var_$n.err; done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/var_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, "ms_per_step", round(d["ms_per_step"], 4), "chunks/s", round(d["value"] / 1e6, 1), "M")
    except Exception as e:
        print(f, "FAILED", e)
PY
python tools/tick_bench.py 2>/dev/null | tail -1 > gpurun_out/tick.json; cat gpurun_out/tick.json
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-strong --segments 0"
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.avg.per_cycle_elapsed --clock-control none -c 70 --csv --log-file gpurun_out/launches.csv $B > /dev/null 2>&1
for n in "$@"; do
  L=$PWD/variants/$n.so; [ "$n" = default ] && L=
  SSE_LIB=$L ncu --set full --import-source on --clock-control none --kernel-name regex:sse_decode_kernel -s 6 -c 2 -o gpurun_out/decode_$n -f $B > /dev/null 2>&1
done
ls gpurun_out | head -40
