# A/B of the build-time variants only (no tests, no profiles): bash tools/ab_quick.sh
mkdir -p gpurun_out
Q="--steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-strong --segments 0"
for v in variants/*.so; do n=$(basename $v .so); SSE_LIB=$PWD/$v python bench.py $Q > gpurun_out/var_$n.json 2> gpurun_out/var_$n.err; done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/var_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, "ms_per_step", round(d["ms_per_step"], 4), "chunks/s", round(d["value"] / 1e6, 1), "M")
    except Exception as e:
        print(f, "FAILED", e)
PY
