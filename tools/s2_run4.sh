mkdir -p gpurun_out
(timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6) > gpurun_out/s2_tests4.log 2>&1
{ python -m benchmarks.calibration --model m --batch 64 --imgsz 640 --batches 12 2>&1 | tail -1
  python -m benchmarks.calibration --model m --batch 64 --imgsz 640 --batches 12 --channels-last 2>&1 | tail -1; } > gpurun_out/s2_calib.log 2>&1
python tools/ci_bench.py 64 > gpurun_out/s2_ci_full2.log 2>&1
cat gpurun_out/s2_tests4.log; cut -c1-600 gpurun_out/s2_calib.log; cut -c1-330 gpurun_out/s2_ci_full2.log
