mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s2_tests5.log 2>&1
tail -4 gpurun_out/s2_tests5.log
echo "=== TMA"; timeout 300 python tools/ci_bench.py 64 2>&1 | cut -c1-330
echo "=== direct"; VSIQ_CI_TMA=0 timeout 300 python tools/ci_bench.py 64 2>&1 | cut -c1-330
