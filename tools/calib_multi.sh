N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 -m benchmarks.calibration --model m --batch 64 --imgsz 640 --batches 48 2>&1 | grep -E '^\{|Error|error' | head -5
