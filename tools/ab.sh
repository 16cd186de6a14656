python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/mb_merge.py 2>&1 | head -5
p() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']))"; }
python bench.py --no-cpu 2>&1 | tail -1 | p C2
