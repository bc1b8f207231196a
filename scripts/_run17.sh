mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k graphed > gpurun_out/r17_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r17_pytest.log | cut -c1-300
for gr in 1 0 1; do timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --graph $gr > gpurun_out/r17_bench_g$gr.json 2> gpurun_out/r17_bench_g$gr.err; python -c "
import json; d=json.load(open('gpurun_out/r17_bench_g$gr.json')); print('graph=$gr', round(d['ms_per_step'],3), round(d['value'],2), 'e2e', round(d['e2e']['value'],2), round(d['e2e'].get('reference_loop_value',0),2), round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d['loss'])" || tail -5 gpurun_out/r17_bench_g$gr.err; done
