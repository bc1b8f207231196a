# round-2 final validation + evidence (the .ncu-rep stays in /tmp on the box: only text summaries travel back)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/r25_pytest.log 2>&1; echo "pytest rc=$?"; grep -v -i warn gpurun_out/r25_pytest.log | tail -8 | cut -c1-300
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r25_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r25_smoke.log
timeout 600 python bench.py > gpurun_out/r25_bench.json 2> gpurun_out/r25_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r25_bench.json')); print(round(d['ms_per_step'],3), round(d['value'],2), 'e2e', round(d['e2e']['value'],2), round(d['e2e'].get('reference_loop_value',0),2), 'frac', round(d['roofline']['frac'],3), 'mfu', round(d['roofline']['step_mfu'],3), d['clocks'], d['loss'], 'ref_cuda', d.get('reference_cuda_eager',{}).get('value'), 'cpu', d.get('cpu_baseline',{}).get('value'))" || tail -5 gpurun_out/r25_bench.err
timeout 300 python bench.py --workload seggpt --precision bf16 --no-cpu-baseline > gpurun_out/r25_bench_seggpt.json 2> gpurun_out/r25_bench_seggpt.err; echo "seggpt rc=$?"; cut -c1-330 gpurun_out/r25_bench_seggpt.json
timeout 600 ncu --set full --clock-control none --import-source on -f -o /tmp/r25_prof python scripts/prof_kernels.py r02 > gpurun_out/r25_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/r25_ncu_full.log
python scripts/ncu_summary.py /tmp/r25_prof.ncu-rep > gpurun_out/r25_ncu_summary.txt 2>&1; wc -l gpurun_out/r25_ncu_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2600 -c 1800 --csv --log-file gpurun_out/r25_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --graph 0 > gpurun_out/r25_bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
python scripts/launch_breakdown.py gpurun_out/r25_launches.csv > gpurun_out/r25_launch_breakdown.txt 2>&1; head -12 gpurun_out/r25_launch_breakdown.txt
gzip -f gpurun_out/r25_launches.csv; du -sh gpurun_out
