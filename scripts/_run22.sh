mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" > gpurun_out/r22_pytest.log 2>&1; echo "pytest attn rc=$?"; grep -v Warn gpurun_out/r22_pytest.log | tail -5 | cut -c1-300
timeout 300 python scripts/time_attn_parts.py 2>&1 | tail -1
timeout 300 python scripts/trace_attn.py > gpurun_out/r22_trace.txt 2>&1; echo "trace rc=$?"; grep -E "^dq phases" gpurun_out/r22_trace.txt | cut -c1-330
timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r22_bench.json 2> gpurun_out/r22_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r22_bench.json')); print(round(d['ms_per_step'],3), round(d['value'],2), 'e2e', round(d['e2e']['value'],2), round(d['e2e'].get('reference_loop_value',0),2), 'frac', round(d['roofline']['frac'],3), d['clocks'], d['loss'])" || tail -5 gpurun_out/r22_bench.err
