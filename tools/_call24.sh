set -x
CMD3="python tools/sweep.py --workload config3 --reads 25000000 --grid r2 --steps 1"
timeout 900 $CMD3 > gpurun_out/r2_c3_plain.jsonl 2> gpurun_out/r2_c3_plain.err && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"kmb_map_reads_mz_kernel|kmb_log_apply" -s 4 -c 2 -o gpurun_out/r2_v11_config3 $CMD3 > gpurun_out/r2_ncu4.log 2>&1
cat gpurun_out/r2_c3_plain.jsonl | cut -c1-500
tail -3 gpurun_out/r2_ncu4.log
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-files --no-oracle"
timeout 600 $CMD > gpurun_out/r2_ncu5_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"kmb_map_reads_kernel|kmb_log_apply" -s 2 -c 2 -o gpurun_out/r2_v11_config2 $CMD > gpurun_out/r2_ncu5.log 2>&1
tail -1 gpurun_out/r2_ncu5_plain.log | cut -c1-200
tail -3 gpurun_out/r2_ncu5.log
