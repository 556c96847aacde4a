# One GPU call: parity of the default build, then A/B of the build-time variants (tools/build_variants.sh) on the C4 bench and on
# the steady-state tick. /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/ab_round.sh'
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -x -q -m gpu ) > gpurun_out/gputests.log 2>&1; grep -E "passed|failed|error" gpurun_out/gputests.log | tail -2
Q="--steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-strong --segments 0"
python bench.py $Q > gpurun_out/var_default.json 2> gpurun_out/var_default.err
python tools/tick_bench.py 2>/dev/null | tail -1 > gpurun_out/tick_default.json
for v in variants/*.so; do n=$(basename $v .so); SSE_LIB=$PWD/$v python bench.py $Q > gpurun_out/var_$n.json 2> gpurun_out/var_$n.err; SSE_LIB=$PWD/$v python tools/tick_bench.py 2>/dev/null | tail -1 > gpurun_out/tick_$n.json; done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/var_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); t = json.load(open(f.replace("var_", "tick_")))
        print(f, "ms_per_step", round(d["ms_per_step"], 4), "chunks/s", round(d["value"] / 1e6, 1), "M", "| tick ms", round(t["ms_per_tick"], 4))
    except Exception as e:
        print(f, "FAILED", e)
PY
