# A/B of channel-innermost kernel variants: tree build vs build_variants/libvsiq_<v>.so (tools/ci_bench.py)
mkdir -p gpurun_out
for v in tree ${VARIANTS:-occ3 occ3u2 b8}; do
  if [ $v = tree ]; then unset VSIQ_LIB; else export VSIQ_LIB=$PWD/build_variants/libvsiq_$v.so; fi
  echo "=== $v"; python tools/ci_bench.py 64 2>&1
done > gpurun_out/s2_ab_ci2.log 2>&1
cut -c1-400 gpurun_out/s2_ab_ci2.log
