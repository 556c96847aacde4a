python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['ms_per_step'])"
python tools/tick_bench.py 2>/dev/null | tail -1
