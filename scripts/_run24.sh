mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention or gemm" > gpurun_out/r24_pytest.log 2>&1; echo "pytest kernels rc=$?"; grep -v Warn gpurun_out/r24_pytest.log | tail -12 | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "tiny" > gpurun_out/r24_pytest2.log 2>&1; echo "pytest model rc=$?"; tail -4 gpurun_out/r24_pytest2.log | cut -c1-300
for lib in "" painter_b200/libpk_fwd_ts1.so; do echo "== PK_LIB=$lib"; PK_LIB=$lib timeout 300 python scripts/time_attn_parts.py 2>&1 | tail -1 | cut -c1-330; done
timeout 300 python scripts/time_gemms.py 2>&1 | grep -E "gelu|sum over" 
timeout 300 python scripts/trace_attn.py > gpurun_out/r24_trace.txt 2>&1; echo "trace rc=$?"; grep -E "^dq phases|^tile  [3-4]" gpurun_out/r24_trace.txt | cut -c1-400
