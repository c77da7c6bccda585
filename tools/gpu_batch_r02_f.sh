#!/bin/bash
mkdir -p gpurun_out
{
echo "== trace"; timeout 300 python tools/trace_cell4.py 2>&1 | tail -8
echo "== pytest quick"; timeout 900 python -m pytest tests/test_tc_path_gpu.py -m gpu -q -x --no-header 2>&1 | tail -4
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/r02f_bench.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'cell us', d['roofline']['us_per_call'], 'frac', d['roofline']['frac'], {k:(v['us']) for k,v in d['kernels'].items()}, d.get('roofline_step',{}).get('frac'), d.get('cpu_baseline',{}).get('value'))
PY
echo "== ncu traffic"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none --replay-mode application --profile-from-start off --csv --log-file gpurun_out/r02f_ksm_traffic.csv python tools/ncu_cell_capture.py 2>&1 | tail -2
python tools/ncu_traffic.py gpurun_out/r02f_ksm_traffic.csv gpurun_out/r02f_ksm_traffic.json
echo "== ncu full (cell4)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cell4_kernel -s 2 -c 1 -o gpurun_out/r02f_cell4 python tools/ncu_cell_capture.py 2>&1 | tail -2
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --eager > /dev/null 2>&1
python tools/ncu_launch_summary.py gpurun_out/r02f_launches.csv | head -14
} > gpurun_out/r02f_main.log 2>&1
cat gpurun_out/r02f_main.log
