mkdir -p gpurun_out/final2
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py 2>gpurun_out/final2/err_C2.log | tail -1 > gpurun_out/final2/bench_c2_1gpu.json
for w in C1 C3 C4; do python bench.py --workload $w --steps 20 --warmup 5 --no-cpu 2>gpurun_out/final2/err_$w.log | tail -1 > gpurun_out/final2/bench_${w}_1gpu.json; done
python tools/sweep.py --workload C2 > gpurun_out/final2/sweep_c2_1gpu.jsonl 2>gpurun_out/final2/err_sweep.log
for f in gpurun_out/final2/bench_*_1gpu.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), d.get('cpu_baseline') and round(d['cpu_baseline']['value']), round(d['roofline']['frac'],4), round(d['roofline']['avg_launch_us'],1), d['roofline']['kernel'][:20])"; done
python -c "
import json
for l in open('gpurun_out/final2/sweep_c2_1gpu.jsonl'):
    d=json.loads(l); print(d['B'],d['T'],d['ms_per_step'],d['frames_per_s'],d['model_tflops'])"
K='regex:gemm_sm100|ffn_fused|csgu|ctc_|merge_weights|relpos|layernorm|vocab|row_dots|conv2d|scale_add'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --launch-skip 250 -c 250 --csv --log-file gpurun_out/final2/launches_c2.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu > /dev/null 2>&1
python tools/ncu_launches.py gpurun_out/final2/launches_c2.csv
