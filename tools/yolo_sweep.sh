python -m benchmarks.yolo_qat --model s --batch 64 --imgsz 640 --steps 8 --channels-last
python -m benchmarks.yolo_qat --model s --batch 64 --imgsz 640 --steps 8 --channels-last --cuda-graph
python -m benchmarks.yolo_qat --model s --batch 64 --imgsz 640 --steps 8 --channels-last --w-bits 4 --a-bits 8 --asym --per-channel --lsq
python -m benchmarks.yolo_qat --model s --batch 64 --imgsz 640 --steps 8 --channels-last --w-bits 4 --a-bits 8 --asym --per-channel --lsq --cuda-graph
