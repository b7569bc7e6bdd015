(KMB_LIB_PATH=$PWD/kmer_mapper_b200/libkmer_mapper_b200_bounds.so timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_v14_pytest_gpu_bounds.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_v14_pytest_gpu_bounds.log); tail -2 gpurun_out/r2_v14_pytest_gpu_bounds.log
CMD3="python tools/sweep.py --workload config3 --reads 25000000 --grid r2 --steps 1"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"kmb_map_reads_mz_kernel" -s 2 -c 1 -o gpurun_out/r2_v14_config3 $CMD3 > gpurun_out/r2_ncu8.log 2>&1
grep -v "^==PROF== Profiling\|^\s*[0-9]*\. " gpurun_out/r2_ncu8.log | tail -2 | cut -c1-200
