(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_t18.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t18.log); tail -2 gpurun_out/r2_t18.log
timeout 400 python tools/sweep.py --grid window --steps 3 > gpurun_out/r2_sweep_applypol.jsonl 2>gpurun_out/r2_sweep_applypol.err; cut -c1-330 gpurun_out/r2_sweep_applypol.jsonl
for w in config4_k15 config3; do
timeout 900 python bench.py --workload $w --no-files --no-e2e --no-cpu-baseline --steps 3 --warmup 2 > gpurun_out/r2_v9_bench_$w.log 2> gpurun_out/r2_v9_bench_$w.err; tail -1 gpurun_out/r2_v9_bench_$w.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$w', d['value'], d['ms_per_step'], 'kernel', r['kernel_ms'], 'apply', r['apply']['ms'], 'frac', r['frac'], r['frac_step'], d['checks']['sample_counts_bit_exact_vs_oracle'])"; tail -2 gpurun_out/r2_v9_bench_$w.err
done
