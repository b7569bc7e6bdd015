(timeout 900 python -m pytest tests/test_gpu_text.py tests/test_gpu_cli.py -m gpu -x -q > gpurun_out/r2_t35.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t35.log); tail -2 gpurun_out/r2_t35.log
S=$(date +%s)
timeout 1200 python bench.py --full-oracle > gpurun_out/r2_v14_bench_config2.log 2> gpurun_out/r2_v14_bench_config2.err; echo "bench rc=$? wall=$(( $(date +%s) - S )) s"
tail -1 gpurun_out/r2_v14_bench_config2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('config2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', r['frac'], r['frac_step'], 'reads/s', {k:v for k,v in d['reads_per_s'].items() if isinstance(v,float)}, [k for k,v in d['checks'].items() if v is False])"
timeout 600 python tools/gz_device_profile.py 8000000 2> gpurun_out/r2_gz_device3.err | grep '"rep": 2' > gpurun_out/r2_gz_device3.jsonl; cut -c1-230 gpurun_out/r2_gz_device3.jsonl
