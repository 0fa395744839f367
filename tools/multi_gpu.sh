N=${1:-2}
PORT=29531
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 50 --warmup 5 --no-cpu 2>&1 | grep -E '^\{|Error|error' 
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>&1 | grep -E '^\{|Error|error'
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT -m benchmarks.yolo_qat --model s --batch 64 --imgsz 640 --steps 8 --w-bits 4 --a-bits 8 --asym --per-channel --lsq 2>&1 | grep -E '^\{|Error|error'
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT -m benchmarks.yolo_qat --model l --batch 16 --imgsz 640 --steps 8 --mixed 2>&1 | grep -E '^\{|Error|error'
