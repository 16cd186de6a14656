p() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']))"; }
python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -2
(cd _ab_old && python bench.py --no-cpu 2>&1 | tail -1 | p old)
python bench.py --no-cpu 2>&1 | tail -1 | p new_default
TAVSR_BRANCH_FORK=0 python bench.py --no-cpu 2>&1 | tail -1 | p new_nofork
