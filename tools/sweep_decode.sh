python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for m in 3 0; do timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e --mode $m 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('mode $m', d['ms_per_step'])"; done
