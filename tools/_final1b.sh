V=$1
S=$(date +%s)
timeout 1200 python bench.py --full-oracle > gpurun_out/r2_${V}_bench_config2.log 2> gpurun_out/r2_${V}_bench_config2.err; echo "bench rc=$? wall=$(( $(date +%s) - S )) s"
tail -1 gpurun_out/r2_${V}_bench_config2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('config2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', r['frac'], r['frac_step'], 'reads/s', {k:v for k,v in d['reads_per_s'].items() if isinstance(v,float)}, d['checks'], d['cpu_baseline'])"
S=$(date +%s)
timeout 1200 python bench.py > gpurun_out/r2_${V}_bench_config2_default.log 2> gpurun_out/r2_${V}_bench_config2_default.err; echo "default bench rc=$? wall=$(( $(date +%s) - S )) s"
