#!/bin/bash
# Re-capture on the FINAL sources after a late library change that does not touch the inference
# kernels: full GPU suite, smoke, the `ncu --set full` digests -> profiles/r02_ncu_traffic.json
# (stamped with the final source digest), then the C2 bench lines that read it.
# usage: gpurun --timeout 1500 -- 'bash tools/final3.sh'
O=gpurun_out/final_r02
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee $O/pytest_gpu.txt
python __graft_entry__.py --smoke 2>&1 | tail -1 | tee $O/smoke.txt
K='regex:gemm_sm100|ffn_fused|csgu|ctc_|merge_|relpos|layernorm|vocab|row_dots|conv2d|scale_add|split_tf32'
for dt in bf16 tf32; do
  ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 345 -c 115 -o $O/ncu_full_c2_$dt python bench.py --steps 2 --warmup 1 --no-graph --no-cpu --no-extra --dtype $dt > /dev/null 2>&1
  ncu -i $O/ncu_full_c2_$dt.ncu-rep --page raw --csv > $O/ncu_full_c2_${dt}_raw.csv 2>/dev/null
  python tools/ncu_summary.py $O/ncu_full_c2_${dt}_raw.csv > $O/ncu_full_c2_${dt}_summary.csv
  head -6 $O/ncu_full_c2_${dt}_summary.csv
done
python tools/ncu_traffic.py bf16=$O/ncu_full_c2_bf16_raw.csv tf32=$O/ncu_full_c2_tf32_raw.csv > $O/r02_ncu_traffic.json
cp $O/r02_ncu_traffic.json profiles/r02_ncu_traffic.json
rm -f $O/*.ncu-rep $O/ncu_full_c2_*_raw.csv
python bench.py --steps 20 --warmup 5 2>$O/err_c2.log | tail -1 > $O/bench_c2_bf16_1gpu.json
python bench.py --steps 20 --warmup 5 --dtype tf32 --no-cpu 2>>$O/err_c2.log | tail -1 > $O/bench_c2_tf32_1gpu.json
for f in $O/bench_c2_bf16_1gpu.json $O/bench_c2_tf32_1gpu.json; do python -c "
import json; d=json.load(open('$f')); r=d['roofline']; t=d.get('train') or {}; print('$f', d['dtype'], round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), round(r['frac'],4), r.get('traffic'), r.get('traffic_build_matches'), (r.get('in_graph') or {}).get('frac'), 'train', t.get('ms_per_step'), (t.get('eager_variant') or {}).get('ms_per_step'))"; done | tee $O/bench_c2_summary.txt
