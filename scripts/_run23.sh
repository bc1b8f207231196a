mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" > gpurun_out/r23_pytest.log 2>&1; echo "pytest attn rc=$?"; grep -v Warn gpurun_out/r23_pytest.log | tail -5 | cut -c1-300
timeout 300 python scripts/time_attn_parts.py 2>&1 | tail -1
timeout 300 python scripts/trace_attn.py > gpurun_out/r23_trace.txt 2>&1; echo "trace rc=$?"; grep -E "^dq phases" gpurun_out/r23_trace.txt | cut -c1-330
