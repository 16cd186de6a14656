#!/bin/bash
# Multi-GPU evidence on ONE box with N GPUs visible (N = first argument, default 8): the C2 bench
# line (inference weak + strong scaling, training step with the overlapped all-reduce), for N = 8
# also C3 and the T x batch sweep.  usage: gpurun --gpus 8 --timeout 900 -- 'bash tools/scale8.sh 8'
N=${1:-8}
O=gpurun_out/scale_r02
mkdir -p $O
run() {  # workload, extra flags, tag
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 --workload $1 $2 2>$O/err_$3.log | tail -1 > $O/bench_$3.json
  python -c "
import json; d=json.load(open('$O/bench_$3.json')); t=d.get('train') or {}
print('$3', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), 'strong', (d.get('strong') or {}).get('value'), 'train', t.get('ms_per_step'), t.get('value'), (t.get('allreduce') or {}), (t.get('eager_variant') or {}).get('ms_per_step'), ((t.get('eager_variant') or {}).get('allreduce') or {}).get('exposed_ms_per_step'))"
}
run C2 "" c2_bf16_${N}gpu
run C2 "--dtype tf32 --no-extra" c2_tf32_${N}gpu
if [ "$N" = "8" ]; then
  run C3 "--no-extra" c3_bf16_8gpu
  run C4 "--no-extra" c4_bf16_8gpu
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
    tools/sweep.py --workload C2 --T 250,500,1000,1500 --batch 1,8,32,128 --steps 8 2>$O/err_sweep8.log > $O/sweep_c2_bf16_8gpu.jsonl
  cat $O/sweep_c2_bf16_8gpu.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['T'], d['batch_per_gpu'], d['ms_per_step'], d['frames_per_s'])"
fi
