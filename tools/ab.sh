python -m pytest tests -m gpu -x -q 2>&1 | tail -4
p() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), d['gpu_launches_per_step'])"; }
python bench.py --no-cpu --workload C1 2>&1 | tail -1 | p C1_im2col
TAVSR_CUDNN_EMBED=1 python bench.py --no-cpu --workload C1 2>&1 | tail -1 | p C1_cudnn
