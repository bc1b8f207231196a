mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" > gpurun_out/r20_pytest.log 2>&1; echo "pytest attn rc=$?"; tail -15 gpurun_out/r20_pytest.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "tiny or graphed" > gpurun_out/r20_pytest2.log 2>&1; echo "pytest model rc=$?"; tail -5 gpurun_out/r20_pytest2.log | cut -c1-300
timeout 300 python scripts/time_attn_parts.py 2>&1 | tail -1
timeout 300 python scripts/trace_attn.py > gpurun_out/r20_trace.txt 2>&1; echo "trace rc=$?"; grep -E "^dq phases|^dkv it  [2-5]|^dq  it  [2-4]" gpurun_out/r20_trace.txt | cut -c1-330
