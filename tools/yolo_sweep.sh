python -m benchmarks.yolo_qat --model n --batch 2 --imgsz 320 --steps 30
python -m benchmarks.yolo_qat --model n --batch 2 --imgsz 320 --steps 30 --cuda-graph
python -m benchmarks.yolo_qat --model s --batch 64 --imgsz 640 --steps 8
python -m benchmarks.yolo_qat --model s --batch 64 --imgsz 640 --steps 8 --cuda-graph
