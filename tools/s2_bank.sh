# weight-bank validation: GPU tests, YOLOv8n/s with and without the bank, bench.py with the size sweep
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15) > gpurun_out/s2_tests2.log 2>&1
{
for extra in "" "--weight-bank" "--cuda-graph" "--cuda-graph --weight-bank"; do
  python -m benchmarks.yolo_qat --model n --batch 2 --imgsz 320 --steps 30 --channels-last $extra 2>&1 | tail -1
done
for extra in "" "--weight-bank"; do
  python -m benchmarks.yolo_qat --model s --batch 64 --imgsz 640 --steps 8 --channels-last --w-bits 4 --a-bits 8 --asym --per-channel --lsq $extra 2>&1 | tail -1
done
} > gpurun_out/s2_bank_yolo.log 2>&1
python bench.py --steps 100 > gpurun_out/s2_bench_sweep.json 2> gpurun_out/s2_bench_sweep.err
cat gpurun_out/s2_tests2.log; cut -c1-700 gpurun_out/s2_bank_yolo.log; tail -3 gpurun_out/s2_bench_sweep.err; cut -c1-300 gpurun_out/s2_bench_sweep.json
