#!/bin/bash
# round-end evidence at HEAD: bench (default + reference arm), launch list, ncu --set full of a layer's kernels, K-SM DRAM traffic
tag=${1:-r02_head}
mkdir -p gpurun_out
{
echo "== bench (default)"; (time python bench.py) > gpurun_out/${tag}_bench_stdout.txt 2>&1; grep "^{" gpurun_out/${tag}_bench_stdout.txt | tail -1 > gpurun_out/${tag}_bench.json; tail -4 gpurun_out/${tag}_bench_stdout.txt | grep real
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'cell us', d['roofline']['us_per_call'], 'frac', d['roofline']['frac'], {k:(v['us']) for k,v in d['kernels'].items()}, 'step frac', d['roofline_step']['frac'])
print(json.dumps(d.get('other_configs')))
print(json.dumps(d.get('train'))[:300])
print(json.dumps(d.get('cpu_baseline'))[:300])
PY
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/${tag}_bench_reference_arm.json; cut -c1-400 gpurun_out/${tag}_bench_reference_arm.json
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_all.csv python tools/cfg_layer_run.py cfg2 12 > gpurun_out/${tag}_launches_run.log 2>&1
python tools/ncu_keep_last.py gpurun_out/${tag}_launches_all.csv 305 > gpurun_out/${tag}_launches.csv   # five timed forwards x (12 x 5 kernels + final LayerNorm)
python tools/ncu_launch_summary.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches_summary.txt; head -12 gpurun_out/${tag}_launches_summary.txt
echo "== ncu traffic"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none --replay-mode application --profile-from-start off --csv --log-file gpurun_out/${tag}_ksm_traffic.csv python tools/ncu_cell_capture.py 2>&1 | tail -1
python tools/ncu_traffic.py gpurun_out/${tag}_ksm_traffic.csv gpurun_out/${tag}_ksm_traffic.json; head -8 gpurun_out/${tag}_ksm_traffic.json
echo "== ncu full (layer)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cell4_kernel|ffn3_kernel|glu4_kernel|conv_kernel" -s 5 -c 5 -o gpurun_out/${tag}_layer -f python tools/prof_modules.py layer 2 2>&1 | tail -1
python tools/ncu_layer_summary.py gpurun_out/${tag}_layer.ncu-rep "ncu --set full --clock-control none --import-source on -k regex:cell4_kernel|ffn3_kernel|glu4_kernel|conv_kernel -s 5 -c 5, python tools/prof_modules.py layer 2: the SECOND call of one ConformerEncoderLayer at the bench shape (B=32 T=1000 D=256, bf16), round-2 HEAD; cold caches, serialised; units in row 3" > gpurun_out/${tag}_ncu_layer_summary.csv
cut -d, -f1,2,3,4,7,9,11 gpurun_out/${tag}_ncu_layer_summary.csv | tail -6
} > gpurun_out/${tag}_main.log 2>&1
cat gpurun_out/${tag}_main.log
