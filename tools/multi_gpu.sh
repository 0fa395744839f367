# usage: bash tools/multi_gpu.sh N [full|lite]  -- the scaling sweep at N GPUs of one node (N = 1 runs without torchrun)
# full: contract bench (both arms) + every YOLO / calibration line; lite: the headline lines only (GPU-minutes are N x).
N=${1:-2}
MODE=${2:-full}
PORT=29531
if [ "$N" = "1" ]; then RUN="python"; RUNM="python -m"; else
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT"; RUNM="$RUN -m"; fi
f() { grep -E '^\{|Error|error' ; }
LSQ="--w-bits 4 --a-bits 8 --asym --per-channel --lsq"
$RUN bench.py --gpus $N --steps 100 --warmup 5 --no-cpu --no-sweep 2>&1 | f
if [ "$MODE" = "full" ]; then
$RUN bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>&1 | f
$RUNM benchmarks.yolo_qat --model s --batch 64 --imgsz 640 --steps 8 --channels-last --weight-bank 2>&1 | f
$RUNM benchmarks.yolo_qat --model l --batch 16 --imgsz 640 --steps 8 --mixed --channels-last --weight-bank 2>&1 | f
fi
$RUNM benchmarks.yolo_qat --model s --batch 64 --imgsz 640 --steps 8 --channels-last --weight-bank $LSQ 2>&1 | f
$RUNM benchmarks.yolo_qat --model s --batch 64 --imgsz 640 --steps 8 --channels-last --weight-bank --cuda-graph $LSQ 2>&1 | f
$RUNM benchmarks.yolo_qat --model l --batch 16 --imgsz 640 --steps 8 --mixed --channels-last --weight-bank --cuda-graph 2>&1 | f
$RUNM benchmarks.calibration --model m --batch 64 --imgsz 640 --batches 48 --channels-last 2>&1 | f
