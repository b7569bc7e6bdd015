set -x
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_t26.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t26.log); tail -2 gpurun_out/r2_t26.log
(KMB_LIB_PATH=$PWD/kmer_mapper_b200/libkmer_mapper_b200_bounds.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "read_table or mz or variant" > gpurun_out/r2_t26b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t26b.log); tail -2 gpurun_out/r2_t26b.log
timeout 900 python bench.py --workload config3 --no-files --no-e2e --no-cpu-baseline --steps 4 --warmup 2 > gpurun_out/r2_v11_config3.log 2> gpurun_out/r2_v11_config3.err; tail -1 gpurun_out/r2_v11_config3.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('config3', d['value'], d['ms_per_step'], 'kernel', r['kernel_ms'], 'apply', r['apply']['ms'], r['frac'], r['frac_step'], r['sector_fetches_per_kmer'], d['checks'])"; tail -2 gpurun_out/r2_v11_config3.err
timeout 600 python bench.py --workload config2 --no-files --no-e2e --no-cpu-baseline --no-oracle --steps 3 --warmup 2 --opt read_table=1 > gpurun_out/r2_v11_config2_rt.log 2> gpurun_out/r2_v11_config2_rt.err; tail -1 gpurun_out/r2_v11_config2_rt.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('config2 read_table', d['value'], d['ms_per_step'], 'kernel', r['kernel_ms'], r['sector_fetches_per_kmer'])"
