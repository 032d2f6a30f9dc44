timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "adversarial or forced_fallback or wide_state" 2>&1 | tail -3
SRUKF_FORCE_FALLBACK_PPM=1000000 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/quick_parity.py 56:2:1 > gpurun_out/r03j_memcheck_seq16.log 2>&1; echo "memcheck seq16 rc=$?"; tail -3 gpurun_out/r03j_memcheck_seq16.log
timeout 300 python bench.py --sweep 66:8192 --steps 3 --warmup 3 --adversarial-frac 0.01 | cut -c1-700
