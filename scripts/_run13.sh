mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" > gpurun_out/r13_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r13_pytest.log
timeout 120 scripts/ubench/softmax_mix > gpurun_out/r13_softmax_mix.txt 2>&1; cat gpurun_out/r13_softmax_mix.txt
timeout 120 scripts/ubench/mma_multi > gpurun_out/r13_mma_multi.txt 2>&1; cat gpurun_out/r13_mma_multi.txt
timeout 300 python scripts/time_attn_parts.py > gpurun_out/r13_attn_parts.txt 2>&1; cat gpurun_out/r13_attn_parts.txt
timeout 300 python scripts/trace_attn.py > gpurun_out/r13_trace.log 2>&1; head -16 gpurun_out/r13_trace.log | cut -c1-330
