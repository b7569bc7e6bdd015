set -x
CMD3="python tools/sweep.py --workload config3 --reads 25000000 --grid r2 --steps 1"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"kmb_map_reads_mz_kernel|kmb_log_apply" -s 2 -c 2 -o gpurun_out/r2_v11_config3 $CMD3 > gpurun_out/r2_ncu4.log 2>&1
grep -v "^==PROF== Profiling\|^\s*[0-9]*\. " gpurun_out/r2_ncu4.log | tail -5
