set -x
export KMB_LIB_PATH=$PWD/kmer_mapper_b200/libkmer_mapper_b200_cta3.so
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --opt async_sectors=1 --opt apply_window_log2=23"
$CMD > gpurun_out/r2_ncu_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"kmb_map_reads_kernel|kmb_log_apply" -s 2 -c 2 -o gpurun_out/r2_async_cta3 $CMD > gpurun_out/r2_ncu.log 2>&1
tail -2 gpurun_out/r2_ncu_plain.log | cut -c1-600
tail -3 gpurun_out/r2_ncu.log
