(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t20.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t20.log); tail -3 gpurun_out/r2_t20.log
for w in config1 config3; do
timeout 900 python bench.py --workload $w --no-files --no-e2e --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r2_v9_bench_$w.log 2> gpurun_out/r2_v9_bench_$w.err; tail -1 gpurun_out/r2_v9_bench_$w.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$w', d['value'], d['ms_per_step'], 'kernel', r['kernel_ms'], 'apply', r['apply'], 'frac', r['frac'], r['frac_step'], d['gpu_launches'], d['checks']['sample_counts_bit_exact_vs_oracle'])"; tail -2 gpurun_out/r2_v9_bench_$w.err
done
