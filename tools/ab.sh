python -m pytest tests -m gpu -x -q 2>&1 | tail -4
p() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), d['roofline']['avg_launch_us'])"; }
python bench.py --no-cpu 2>&1 | tail -1 | p rowwarp
TAVSR_DEBUG=11=1 python bench.py --no-cpu 2>&1 | tail -1 | p threadrow
python bench.py --no-cpu --batch 8 2>&1 | tail -1 | p rowwarp_b8
TAVSR_DEBUG=11=1 python bench.py --no-cpu --batch 8 2>&1 | tail -1 | p threadrow_b8
