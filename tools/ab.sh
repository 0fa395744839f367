python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -3
echo "=== base"; python tools/microbench.py --log2n 24 26 28 --iters 10 --per-channel 2>&1 | grep -E "copy_|fq_fwd   W8 host|lsq_bwd|ste_bwd|fused|observe"
for v in u4; do echo "=== $v"; VSIQ_LIB=$PWD/build_variants/libvsiq_$v.so python tools/microbench.py --log2n 26 28 --iters 10 2>&1 | grep -E "fq_fwd   W8 host|lsq_bwd|ste_bwd|fused|observe"; done
