# Round-end evidence on ONE B200: tests, contract bench (both arms), ncu launch lists + full captures, microbenchmarks.
# Everything lands in gpurun_out/ with the tag given as $1 (default r02); tools/collect_profiles.sh copies the summaries
# into profiles/.  Nothing printed under ncu is ever used as a benchmark number.  The .ncu-rep files stay in /tmp on the
# box (gpurun brings back at most 64 MiB); only their raw CSV exports travel.
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; tail -2 $O/${TAG}_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_native.json 2> $O/${TAG}_bench_native.err; tail -c 300 $O/${TAG}_bench_native.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/${TAG}_bench_reference.json 2>/dev/null
MICRO="--steps 20 --warmup 3 --no-e2e --no-cpu --no-sweep --no-yolo --no-calibration --no-gpu-eager"
python bench.py $MICRO > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/${TAG}_launches_bench.csv \
    python bench.py $MICRO > $O/${TAG}_ncu_list.log 2>&1
python tools/ncu_driver.py 28 2 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"fq_|lsq_|observe|ci_|mt_" -s 19 -c 19 -o /tmp/${TAG}_full \
    python tools/ncu_driver.py 28 2 > $O/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > $O/${TAG}_ncu_full_raw.csv 2>/dev/null
for s in "512 20" "256 40" "128 80"; do
  python tools/ci_ncu_driver.py $s > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_ci_${s// /x}.csv \
      python tools/ci_ncu_driver.py $s > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:"ci_bwd_tma|ci_fwd|ci_observe_kernel|ci_epilogue_observe" -s 6 -c 6 \
    -o /tmp/${TAG}_full_ci_256x40 python tools/ci_ncu_driver.py 256 40 2 > /dev/null 2>&1
ncu -i /tmp/${TAG}_full_ci_256x40.ncu-rep --page raw --csv > $O/${TAG}_ncu_full_ci_256x40_raw.csv 2>/dev/null
python tools/microbench.py --log2n 22 24 26 28 --iters 10 --per-channel > $O/${TAG}_microbench.log 2>&1
python tools/ci_bench.py 64 > $O/${TAG}_ci_bench.log 2>&1
python tools/shapes_bench.py 64 > $O/${TAG}_shapes_bench.log 2>&1
LSQ="--w-bits 4 --a-bits 8 --asym --per-channel --lsq"
{
python -m benchmarks.yolo_qat --model n --batch 2 --imgsz 320 --steps 30 --channels-last --weight-bank
python -m benchmarks.yolo_qat --model n --batch 2 --imgsz 320 --steps 30 --channels-last --weight-bank --cuda-graph
python -m benchmarks.yolo_qat --model n --batch 2 --imgsz 320 --steps 30 --channels-last --quant-impl eager
python -m benchmarks.yolo_qat --model s --batch 64 --imgsz 640 --steps 16 --channels-last --weight-bank --cuda-graph
python -m benchmarks.yolo_qat --model s --batch 64 --imgsz 640 --steps 16 --channels-last --weight-bank $LSQ --profile-steps 2
python -m benchmarks.yolo_qat --model s --batch 64 --imgsz 640 --steps 16 --channels-last --weight-bank --cuda-graph $LSQ --profile-steps 2
python -m benchmarks.yolo_qat --model s --batch 64 --imgsz 640 --steps 8 --channels-last --quant-impl eager $LSQ
python -m benchmarks.yolo_qat --model l --batch 16 --imgsz 640 --steps 8 --channels-last --weight-bank --cuda-graph --mixed
} 2>&1 | grep -E '^\{|Error|error' > $O/${TAG}_yolo_1gpu.log
python tools/step_profile.py --channels-last --weight-bank $LSQ > $O/${TAG}_step_profile.log 2>&1
python -m benchmarks.calibration --model m --batch 64 --imgsz 640 --batches 50 --channels-last 2>&1 | grep -E '^\{' > $O/${TAG}_calibration_1gpu.log
cut -c1-200 $O/${TAG}_yolo_1gpu.log; cut -c1-400 $O/${TAG}_calibration_1gpu.log; head -3 $O/${TAG}_step_profile.log
