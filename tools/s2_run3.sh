mkdir -p gpurun_out
(timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6) > gpurun_out/s2_tests3.log 2>&1
python tools/ci_bench.py 64 > gpurun_out/s2_ci_full.log 2>&1
python tools/step_profile.py --channels-last --w-bits 4 --a-bits 8 --asym --per-channel --lsq > gpurun_out/s2_step_profile.log 2>&1
cat gpurun_out/s2_tests3.log; cut -c1-330 gpurun_out/s2_ci_full.log; cut -c1-200 gpurun_out/s2_step_profile.log
