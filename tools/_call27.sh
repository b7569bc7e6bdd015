set -x
timeout 900 python tools/sweep.py --workload config3 --reads 25000000 --grid filt --steps 2 > gpurun_out/r2_c3_filt.jsonl 2> gpurun_out/r2_c3_filt.err; cut -c1-420 gpurun_out/r2_c3_filt.jsonl; tail -2 gpurun_out/r2_c3_filt.err
CMD3="python tools/sweep.py --workload config3 --reads 25000000 --grid r2 --steps 1"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"kmb_map_reads_mz_kernel" -s 2 -c 1 -o gpurun_out/r2_v12_config3 $CMD3 > gpurun_out/r2_ncu6.log 2>&1
grep -v "^==PROF== Profiling\|^\s*[0-9]*\. " gpurun_out/r2_ncu6.log | tail -3 | cut -c1-300
