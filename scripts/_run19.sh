# graphed-step test failure, TS-form ubench, forward TS vs SS timing, attention kernel tests, traces
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu > gpurun_out/r19_pytest_model.log 2>&1; echo "pytest model rc=$?"; grep -v Warning gpurun_out/r19_pytest_model.log | tail -40 | cut -c1-400
timeout 120 scripts/ubench/mma_ts > gpurun_out/r19_mma_ts.txt 2>&1; echo "mma_ts rc=$?"; cat gpurun_out/r19_mma_ts.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" > gpurun_out/r19_pytest.log 2>&1; echo "pytest attn rc=$?"; tail -5 gpurun_out/r19_pytest.log | cut -c1-300
for lib in "" painter_b200/libpk_fwd_ss.so; do echo "== PK_LIB=$lib"; PK_LIB=$lib timeout 300 python scripts/time_attn_parts.py 2>&1 | tail -1; done
timeout 300 python scripts/trace_attn.py > gpurun_out/r19_trace.txt 2>&1; echo "trace rc=$?"; cat gpurun_out/r19_trace.txt
du -sh gpurun_out
