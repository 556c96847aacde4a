# GPU check of the current build (run on the GPU box via gpurun).
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for m in 3 0; do
  timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --mode $m 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('mode $m', d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['zero_copy_frames'], d['e2e'])"
done
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --mode 3 --flags 8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('copy_out', d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e'])"
