timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -k "outside or writes_only or long_tiles or 2p30 or code_export" > gpurun_out/r02k_newtests.log 2>&1; tail -4 gpurun_out/r02k_newtests.log
python tools/codes_bench.py > gpurun_out/r02k_codes_bench.log 2>&1; grep "n=2" gpurun_out/r02k_codes_bench.log
python tools/codes_bench.py --once > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fq_codes -s 1 -c 1 -o /tmp/codes8 python tools/codes_bench.py --once > gpurun_out/r02k_ncu_codes.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fq_codes -s 3 -c 1 -o /tmp/codes4 python tools/codes_bench.py --once >> gpurun_out/r02k_ncu_codes.log 2>&1
ncu -i /tmp/codes8.ncu-rep --page raw --csv > gpurun_out/r02k_ncu_codes_int8_raw.csv 2>/dev/null
ncu -i /tmp/codes4.ncu-rep --page raw --csv > gpurun_out/r02k_ncu_codes_int4_raw.csv 2>/dev/null
tail -3 gpurun_out/r02k_ncu_codes.log
