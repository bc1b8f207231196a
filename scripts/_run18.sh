# round-2 validation + evidence: full GPU suite, smoke, default bench, ncu launch list of the bench, ncu --set full of the hot kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
( time timeout 1500 python -m pytest tests -x -q -m gpu --durations=15 ) > gpurun_out/r18_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r18_pytest.log | cut -c1-200
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r18_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r18_smoke.log
timeout 600 python bench.py > gpurun_out/r18_bench.json 2> gpurun_out/r18_bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r18_bench.json; tail -3 gpurun_out/r18_bench.err
PK_PROF_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -f -o gpurun_out/r18_prof python scripts/prof_kernels.py gemm attn stream > gpurun_out/r18_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/r18_ncu_full.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r18_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --graph 0 > gpurun_out/r18_bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
ls -la gpurun_out | head -30
