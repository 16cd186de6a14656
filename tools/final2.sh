#!/bin/bash
# Round-2 evidence pass on ONE GPU box: tests, smoke, ncu launch lists + one `ncu --set full` capture
# per compute mode of an eager C2 step (-> profiles/r02_ncu_traffic.json for the bench lines that
# follow, stamped with this build's digest), bench lines of every workload in both modes, the
# training legs, the training launch list, in-kernel phase stamps.
# usage: gpurun --timeout 2700 -- 'bash tools/final2.sh'
O=gpurun_out/final_r02
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee $O/pytest_gpu.txt
python __graft_entry__.py --smoke 2>&1 | tail -1 | tee $O/smoke.txt
K='regex:gemm_sm100|ffn_fused|csgu|ctc_|merge_|relpos|layernorm|vocab|row_dots|conv2d|scale_add|split_tf32'
for dt in bf16 tf32; do
  ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --launch-skip 230 -c 230 --csv --log-file $O/launches_c2_$dt.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu --no-extra --dtype $dt > /dev/null 2>&1
  python tools/ncu_launches.py $O/launches_c2_$dt.csv > $O/launches_c2_${dt}_summary.txt
  head -12 $O/launches_c2_${dt}_summary.txt
  ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 345 -c 115 -o $O/ncu_full_c2_$dt python bench.py --steps 2 --warmup 1 --no-graph --no-cpu --no-extra --dtype $dt > /dev/null 2>&1
  ncu -i $O/ncu_full_c2_$dt.ncu-rep --page raw --csv > $O/ncu_full_c2_${dt}_raw.csv 2>/dev/null
  python tools/ncu_summary.py $O/ncu_full_c2_${dt}_raw.csv > $O/ncu_full_c2_${dt}_summary.csv
  head -10 $O/ncu_full_c2_${dt}_summary.csv
done
python tools/ncu_traffic.py bf16=$O/ncu_full_c2_bf16_raw.csv tf32=$O/ncu_full_c2_tf32_raw.csv > $O/r02_ncu_traffic.json
cp $O/r02_ncu_traffic.json profiles/r02_ncu_traffic.json      # read by the bench runs below
rm -f $O/*.ncu-rep $O/ncu_full_c2_*_raw.csv                   # the digests are what is read back
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/train_launches.csv python tools/prof_train.py > $O/prof_train.log 2>&1
python tools/ncu_launches.py $O/train_launches.csv > $O/train_launches_summary.txt
python tools/prof_train.py 2>&1 | tail -1 > $O/prof_train_dropout.txt
python tools/prof_train.py --eval 2>&1 | tail -1 > $O/prof_train_eval.txt
python tools/time_ffn.py > $O/time_ffn.txt 2>&1
python tools/time_attn.py > $O/time_attn.txt 2>&1
python bench.py --steps 20 --warmup 5 2>$O/err_c2.log | tail -1 > $O/bench_c2_bf16_1gpu.json
python bench.py --steps 20 --warmup 5 --dtype tf32 --no-cpu 2>>$O/err_c2.log | tail -1 > $O/bench_c2_tf32_1gpu.json
python bench.py --steps 5 --warmup 2 --impl reference 2>>$O/err_c2.log | tail -1 > $O/bench_c2_reference_arm.json
for w in C1 C3 C4; do for dt in bf16 tf32; do
  python bench.py --workload $w --steps 20 --warmup 5 --no-cpu --no-extra --dtype $dt 2>$O/err_$w.log | tail -1 > $O/bench_${w}_${dt}_1gpu.json
done; done
for w in C1 C3 C4; do
  python bench.py --workload $w --mode train --train-steps 5 --steps 5 --warmup 3 --no-cpu 2>>$O/err_$w.log | tail -1 > $O/bench_${w}_train_1gpu.json
done
for f in $O/bench_*_bf16_1gpu.json $O/bench_*_tf32_1gpu.json; do python -c "
import json; d=json.load(open('$f')); r=d['roofline']; print('$f', d['dtype'], round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), round(r['frac'],4), round(r['avg_launch_us'],1), r['kernel'][:24], r.get('traffic'), r.get('traffic_build_matches'), (r.get('in_graph') or {}).get('frac'))"; done | tee $O/bench_summary.txt
for f in $O/bench_c2_bf16_1gpu.json $O/bench_C*_train_1gpu.json; do python -c "
import json; d=json.load(open('$f')); t=d.get('train') or {}; e=t.get('eager_variant') or {}
print('$f', 'train graph ms', t.get('ms_per_step'), 'frames/s', t.get('value'), 'eager ms', e.get('ms_per_step'), 'eval-mode ms', e.get('eval_mode_ms_per_step'))"; done | tee $O/train_summary.txt
