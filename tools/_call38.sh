(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_v15_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_v15_pytest_gpu.log); tail -2 gpurun_out/r2_v15_pytest_gpu.log
(KMB_LIB_PATH=$PWD/kmer_mapper_b200/libkmer_mapper_b200_bounds.so timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_v15_pytest_gpu_bounds.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_v15_pytest_gpu_bounds.log); tail -2 gpurun_out/r2_v15_pytest_gpu_bounds.log
timeout 600 python tools/sweep.py --workload config3 --reads 50000000 --grid window3 --steps 3 2> gpurun_out/r2_c3_win.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['opts'], 'kernel_ms', round(d['kernel_ms'],3), 'apply_ms', round(d['apply_ms'],3), 'step_ms', round(d['step_ms'],3), d['counts_equal_first'])"
timeout 1500 python bench.py --workload config3 --no-files --no-cpu-baseline --no-e2e --steps 4 --warmup 3 > gpurun_out/r2_v15_bench_config3.log 2> gpurun_out/r2_v15_bench_config3.err; echo "config3 rc=$?"; tail -1 gpurun_out/r2_v15_bench_config3.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('config3', d['value'], d['ms_per_step'], 'kernel', r['kernel_ms'], 'apply', r['apply']['ms'], 'frac', r['frac'], r['frac_step'], 'nonkernel', 1-r['kernel_share_of_step'], {k:v for k,v in d['checks'].items() if 'oracle' in k})"
timeout 600 python bench.py --workload config2 --no-files --no-e2e --no-cpu-baseline --no-oracle --steps 3 --warmup 2 --opt read_table=1 2> /dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('config2 read_table', d['value'], d['ms_per_step'], 'kernel', r['kernel_ms'], 'apply', r['apply']['ms'])"
