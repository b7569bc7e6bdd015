CMD="python bench.py --workload config4_k15 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-files --no-oracle"
timeout 600 $CMD > gpurun_out/r2_k15_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"kmb_map_reads_kernel|kmb_log_apply" -s 2 -c 2 -o gpurun_out/r2_v12_k15 $CMD > gpurun_out/r2_ncu7.log 2>&1
tail -1 gpurun_out/r2_k15_plain.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('k15', d['value'], d['ms_per_step'], 'kernel', r['kernel_ms'], 'apply', r['apply']['ms'])"
tail -2 gpurun_out/r2_ncu7.log | cut -c1-200
timeout 900 python tools/api_throughput.py > gpurun_out/r2_api_throughput.jsonl 2> gpurun_out/r2_api_throughput.err; cat gpurun_out/r2_api_throughput.jsonl | cut -c1-400; tail -3 gpurun_out/r2_api_throughput.err
