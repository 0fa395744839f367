python -m benchmarks.yolo_qat --model s --batch 64 --imgsz 640 --steps 8 --channels-last
python -m benchmarks.yolo_qat --model s --batch 64 --imgsz 640 --steps 8 --channels-last --cuda-graph
python -m benchmarks.yolo_qat --model s --batch 64 --imgsz 640 --steps 8 --channels-last --quant-impl eager
python -m benchmarks.yolo_qat --model n --batch 2 --imgsz 320 --steps 30 --channels-last --cuda-graph
