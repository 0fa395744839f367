LSQ="--w-bits 4 --a-bits 8 --asym --per-channel --lsq"
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -2
for extra in "--no-prefetch" "" "--cuda-graph --no-prefetch" "--cuda-graph"; do
python -m benchmarks.yolo_qat --model s --batch 64 --imgsz 640 --steps 16 --channels-last --weight-bank $LSQ $extra 2>&1 | grep -E '^\{|Error|error' | cut -c1-420
done
python tools/ci_bench.py 64 2>&1 | cut -c290-400
