python tools/pitch_count.py --w-bits 4 --a-bits 8 --asym --per-channel --lsq 2>&1 | tail -40
python tools/ci_bench.py 64 2>&1 | cut -c180-420
for extra in "" "--weight-bank"; do
python -m benchmarks.yolo_qat --model s --batch 64 --imgsz 640 --steps 8 --channels-last --w-bits 4 --a-bits 8 --asym --per-channel --lsq $extra 2>&1 | tail -1 | cut -c1-1200
done
