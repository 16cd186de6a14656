mkdir -p gpurun_out/final
for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 2>gpurun_out/final/err_c2_$n.log | tail -1 > gpurun_out/final/bench_c2_${n}gpu.json
python -c "
import json; d=json.load(open('gpurun_out/final/bench_c2_${n}gpu.json')); print('C2 x$n', round(d['value']), d['ms_per_step'], round(d['e2e']['value']))"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --steps 20 --warmup 5 --workload C3 2>gpurun_out/final/err_c3_8.log | tail -1 > gpurun_out/final/bench_c3_8gpu.json
python -c "
import json; d=json.load(open('gpurun_out/final/bench_c3_8gpu.json')); print('C3 x8', round(d['value']), d['ms_per_step'], round(d['e2e']['value']))"
