# Short round-end check on ONE B200 (the long form is tools/final_1gpu.sh): GPU tests, smoke(), both contract-bench arms
# and the ncu launch list of the microbench step.  Outputs -> gpurun_out/<tag>_*.
TAG=${1:-r02n}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; tail -2 $O/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_native.json 2> $O/${TAG}_bench_native.err; tail -c 300 $O/${TAG}_bench_native.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/${TAG}_bench_reference.json 2>/dev/null
MICRO="--steps 20 --warmup 3 --no-e2e --no-cpu --no-sweep --no-yolo --no-calibration --no-gpu-eager"
python bench.py $MICRO > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/${TAG}_launches_bench.csv \
    python bench.py $MICRO > $O/${TAG}_ncu_list.log 2>&1
cut -c1-400 $O/${TAG}_bench_native.json; cut -c1-300 $O/${TAG}_bench_reference.json
