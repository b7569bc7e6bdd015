(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_t22.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t22.log); tail -2 gpurun_out/r2_t22.log
for w in config1 config3 config2; do
timeout 900 python bench.py --workload $w --no-files --no-e2e --no-cpu-baseline --no-oracle --steps 4 --warmup 2 > gpurun_out/r2_v10_$w.log 2> gpurun_out/r2_v10_$w.err; tail -1 gpurun_out/r2_v10_$w.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$w', d['value'], d['ms_per_step'], 'kernel', r['kernel_ms'], 'apply', r['apply']['ms'], r['frac'], r['frac_step'])"; tail -2 gpurun_out/r2_v10_$w.err
done
