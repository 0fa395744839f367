# A/B of the channel-innermost kernels: tree build (3 CTAs/SM, shared-memory fp64 sums) vs build_variants/libvsiq_head.so
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5) > gpurun_out/s2_tests.log 2>&1
for v in tree head; do
  if [ $v = head ]; then export VSIQ_LIB=$PWD/build_variants/libvsiq_head.so; else unset VSIQ_LIB; fi
  echo "=== $v"; python tools/ci_bench.py 64 2>&1
  python -m benchmarks.yolo_qat --model s --batch 64 --imgsz 640 --steps 8 --channels-last --w-bits 4 --a-bits 8 --asym --per-channel --lsq 2>&1 | tail -1
done > gpurun_out/s2_ab_ci.log 2>&1
cat gpurun_out/s2_tests.log; cat gpurun_out/s2_ab_ci.log | cut -c1-400
