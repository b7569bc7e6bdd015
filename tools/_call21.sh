for o in 23 24; do
timeout 900 python bench.py --workload config3 --no-files --no-e2e --no-cpu-baseline --no-oracle --steps 4 --warmup 2 --opt apply_window_log2=$o > gpurun_out/r2_c3_w$o.log 2> gpurun_out/r2_c3_w$o.err; tail -1 gpurun_out/r2_c3_w$o.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('window $o', d['value'], d['ms_per_step'], 'kernel', r['kernel_ms'], 'apply', r['apply']['ms'])"; tail -2 gpurun_out/r2_c3_w$o.err
done
