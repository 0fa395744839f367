RUNM="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 -m"
LSQ="--w-bits 4 --a-bits 8 --asym --per-channel --lsq"
$RUNM benchmarks.yolo_qat --model s --batch 64 --imgsz 640 --steps 8 --channels-last --weight-bank $LSQ 2>&1 | grep -E '^\{|Error|error' | cut -c1-330
timeout 300 $RUNM benchmarks.yolo_qat --model s --batch 64 --imgsz 640 --steps 8 --channels-last --weight-bank --cuda-graph $LSQ 2>&1 | grep -E '^\{|Error|error|Traceback' | cut -c1-330
timeout 300 $RUNM benchmarks.yolo_qat --model s --batch 64 --imgsz 640 --steps 8 --channels-last --cuda-graph $LSQ 2>&1 | grep -E '^\{|Error|error|Traceback' | cut -c1-330
