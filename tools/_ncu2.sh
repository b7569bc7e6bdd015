set -x
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-files --no-oracle"
$CMD > gpurun_out/r2_ncu2_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"kmb_map_reads_kernel|kmb_log_apply" -s 2 -c 2 -o gpurun_out/r2_v8_config2 $CMD > gpurun_out/r2_ncu2.log 2>&1
tail -1 gpurun_out/r2_ncu2_plain.log | cut -c1-300
tail -3 gpurun_out/r2_ncu2.log
