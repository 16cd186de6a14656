mkdir -p gpurun_out/final
python bench.py 2>gpurun_out/final/err_C2.log | tail -1 > gpurun_out/final/bench_c2_1gpu.json
python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tail -1 > gpurun_out/final/bench_c2_reference_arm.json
for w in C1 C3 C4; do python bench.py --workload $w --steps 20 --warmup 5 2>gpurun_out/final/err_$w.log | tail -1 > gpurun_out/final/bench_${w}_1gpu.json; done
python tools/sweep.py --workload C2 > gpurun_out/final/sweep_c2_1gpu.jsonl 2>gpurun_out/final/err_sweep.log
python tools/sweep.py --workload C4 --T 250,500,1000 --batch 1,8,32,128 > gpurun_out/final/sweep_c4_1gpu.jsonl 2>>gpurun_out/final/err_sweep.log
python tools/sweep.py --workload C1 --T 249 --batch 1,8,32,128 > gpurun_out/final/sweep_c1_1gpu.jsonl 2>>gpurun_out/final/err_sweep.log
for f in gpurun_out/final/bench_*_1gpu.json gpurun_out/final/bench_c2_reference_arm.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), d.get('cpu_baseline') and round(d['cpu_baseline']['value']), d.get('roofline',{}).get('frac'))"; done
cat gpurun_out/final/sweep_c4_1gpu.jsonl gpurun_out/final/sweep_c1_1gpu.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['workload'],d['B'],d['T'],d['ms_per_step'],d['frames_per_s'],d['model_tflops'])"
tail -c 400 gpurun_out/final/err_sweep.log
