mkdir -p gpurun_out
for w in C2 C3; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 --workload $w 2>gpurun_out/err8_$w.log | tail -1 > gpurun_out/bench8_$w.json
tail -c 300 gpurun_out/err8_$w.log
python -c "
import json; d=json.load(open('gpurun_out/bench8_$w.json')); print('$w x8', round(d['value']), d['ms_per_step'], round(d['e2e']['value']))"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench2_C2.json
python -c "
import json; d=json.load(open('gpurun_out/bench2_C2.json')); print('C2 x2', round(d['value']), d['ms_per_step'], round(d['e2e']['value']))"
