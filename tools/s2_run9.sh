RUNM="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 -m"
LSQ="--w-bits 4 --a-bits 8 --asym --per-channel --lsq"
$RUNM benchmarks.yolo_qat --model s --batch 64 --imgsz 640 --steps 16 --channels-last --weight-bank --cuda-graph $LSQ 2>&1 | grep -E '^\{|Error|error'
$RUNM benchmarks.yolo_qat --model l --batch 16 --imgsz 640 --steps 16 --mixed --channels-last --weight-bank --cuda-graph 2>&1 | grep -E '^\{|Error|error'
