(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t17.log); tail -3 gpurun_out/r2_t17.log
timeout 900 python bench.py --full-oracle > gpurun_out/r2_v8_bench_config2.log 2> gpurun_out/r2_v8_bench_config2.err; tail -1 gpurun_out/r2_v8_bench_config2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('c2', d['value'], d['ms_per_step'], json.dumps(d['e2e']), json.dumps(d['reads_per_s']), d['checks'])"; tail -2 gpurun_out/r2_v8_bench_config2.err
timeout 600 python bench.py --impl reference > gpurun_out/r2_v8_bench_config2_reference_arm.log 2>&1; tail -1 gpurun_out/r2_v8_bench_config2_reference_arm.log | cut -c1-400
for w in config1 config4_k21 config4_k15 config5; do
timeout 900 python bench.py --workload $w --no-files > gpurun_out/r2_v8_bench_$w.log 2> gpurun_out/r2_v8_bench_$w.err; tail -1 gpurun_out/r2_v8_bench_$w.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$w', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'] if d['e2e'] else None, 'kernel', r['kernel_ms'], 'apply', r['apply']['ms'], 'frac', r['frac'], r['frac_step'], d['checks'])"; tail -2 gpurun_out/r2_v8_bench_$w.err
done
