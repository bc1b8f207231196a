mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" > gpurun_out/r14_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r14_pytest.log
for lib in libpainter_b200 libpk_fwd_old libpk_fwd_half; do echo "== $lib"; PK_LIB=$PWD/painter_b200/$lib.so timeout 300 python scripts/time_attn_parts.py 2>&1 | tail -1; done | tee gpurun_out/r14_fwd_variants.txt
timeout 300 python scripts/trace_attn.py > gpurun_out/r14_trace.log 2>&1; head -14 gpurun_out/r14_trace.log | cut -c1-300; grep "^dkv" gpurun_out/r14_trace.log | cut -c90-400
timeout 600 python scripts/step_timeline.py 3 > gpurun_out/r14_step_timeline.txt 2> gpurun_out/r14_step_timeline.err; head -60 gpurun_out/r14_step_timeline.txt; tail -5 gpurun_out/r14_step_timeline.err
