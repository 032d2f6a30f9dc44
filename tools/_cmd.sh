tools/ab_bench.sh r03g "kc32|kc32|" "kc16|kc16|" 2>&1 | grep -v QUICK
echo "== sanitizer memcheck (small cases, fused path + forced fallback + wide map)"
SRUKF_FORCE_FALLBACK_PPM=500000 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/quick_parity.py 3:4:2 8:4:2 > gpurun_out/r03g_memcheck_small.log 2>&1; echo "memcheck small rc=$?"; tail -3 gpurun_out/r03g_memcheck_small.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/quick_parity.py 108:2:1 > gpurun_out/r03g_memcheck_wide.log 2>&1; echo "memcheck wide rc=$?"; tail -3 gpurun_out/r03g_memcheck_wide.log
echo "== sanitizer racecheck (forced fallback kernel)"
SRUKF_FORCE_FALLBACK_PPM=1000000 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/quick_parity.py 5:2:1 > gpurun_out/r03g_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r03g_racecheck.log
