mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" > gpurun_out/r21_pytest.log 2>&1; echo "pytest attn rc=$?"; grep -v Warn gpurun_out/r21_pytest.log | tail -25 | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "tiny or graphed" > gpurun_out/r21_pytest2.log 2>&1; echo "pytest model rc=$?"; tail -8 gpurun_out/r21_pytest2.log | cut -c1-300
timeout 300 python scripts/time_attn_parts.py 2>&1 | tail -1
timeout 300 python scripts/trace_attn.py > gpurun_out/r21_trace.txt 2>&1; echo "trace rc=$?"; grep -E "^dq phases|^dkv it  [2-4]|^dq  it  [2-3]" gpurun_out/r21_trace.txt | cut -c1-330
