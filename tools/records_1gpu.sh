# final single-GPU records of the round: tests, smoke, default bench (both arms), the other configs, launch list
set -x
V=$1
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_${V}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_${V}_pytest_gpu.log); tail -2 gpurun_out/r2_${V}_pytest_gpu.log
(timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_${V}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_${V}_smoke.log); tail -2 gpurun_out/r2_${V}_smoke.log
timeout 900 python bench.py --impl reference > gpurun_out/r2_${V}_bench_config2_reference_arm.log 2> gpurun_out/r2_${V}_bench_config2_reference_arm.err; tail -1 gpurun_out/r2_${V}_bench_config2_reference_arm.log | cut -c1-200
timeout 1200 python bench.py --full-oracle > gpurun_out/r2_${V}_bench_config2.log 2> gpurun_out/r2_${V}_bench_config2.err; echo "bench rc=$?"; tail -1 gpurun_out/r2_${V}_bench_config2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('config2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', r['frac'], r['frac_step'], 'reads/s', {k:v for k,v in d['reads_per_s'].items() if isinstance(v,float)}, d['checks'])"
for w in config1 config3 config4_k21 config4_k15 config5; do
timeout 1500 python bench.py --workload $w --no-files --no-cpu-baseline --full-oracle --steps 4 --warmup 3 > gpurun_out/r2_${V}_bench_$w.log 2> gpurun_out/r2_${V}_bench_$w.err; echo "$w rc=$?"; tail -1 gpurun_out/r2_${V}_bench_$w.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$w', d['value'], d['ms_per_step'], 'e2e', (d['e2e'] or {}).get('value'), 'kernel', r['kernel_ms'], 'apply', r['apply']['ms'], 'frac', r['frac'], r['frac_step'], {k:v for k,v in d['checks'].items() if 'oracle' in k})"
done
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-files --no-oracle"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:kmb_ -c 400 --csv --log-file gpurun_out/r2_${V}_config2_launches.csv $CMD > gpurun_out/r2_${V}_ncu_launch.log 2>&1; tail -1 gpurun_out/r2_${V}_ncu_launch.log | cut -c1-100
CMD3="python bench.py --workload config3 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-files --no-oracle"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:kmb_ -c 400 --csv --log-file gpurun_out/r2_${V}_config3_launches.csv $CMD3 > gpurun_out/r2_${V}_ncu_launch3.log 2>&1; tail -1 gpurun_out/r2_${V}_ncu_launch3.log | cut -c1-100
