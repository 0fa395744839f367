# Copies the round-end evidence written by tools/final_1gpu.sh (gpurun_out/<tag>_*) into profiles/ (tracked), condensing
# the ncu raw exports into per-launch summaries.  Runs here, without a GPU.     bash tools/collect_profiles.sh [tag]
TAG=${1:-r02}
O=gpurun_out
P=profiles
for f in pytest_gpu.log bench_native.json bench_reference.json launches_bench.csv launches_ci_512x20.csv \
         launches_ci_256x40.csv launches_ci_128x80.csv microbench.log ci_bench.log shapes_bench.log yolo_1gpu.log \
         step_profile.log calibration_1gpu.log; do
  [ -s $O/${TAG}_$f ] && cp $O/${TAG}_$f $P/${TAG}_$f
done
[ -s $O/${TAG}_ncu_full_raw.csv ] && python tools/ncu_summary.py $O/${TAG}_ncu_full_raw.csv $P/${TAG}_ncu_full_summary.csv \
    --traffic-tag "@${TAG}:tools/ncu_driver.py 28 2" --overwrite
[ -s $O/${TAG}_ncu_full_ci_256x40_raw.csv ] && python tools/ncu_summary.py $O/${TAG}_ncu_full_ci_256x40_raw.csv \
    $P/${TAG}_ncu_full_ci_256x40_summary.csv --traffic-tag "@${TAG}:[64,256,40,40]" --overwrite
python tools/sass_excerpt.py > $P/${TAG}_sass_excerpt.txt
ls -la $P/${TAG}_* | wc -l
