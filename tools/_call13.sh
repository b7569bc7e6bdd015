set -x
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t13.log); tail -5 gpurun_out/r2_t13.log
for v in "" _bins8 _unroll1; do
KMB_LIB_PATH=$PWD/kmer_mapper_b200/libkmer_mapper_b200$v.so timeout 400 python tools/sweep.py --grid r2 --steps 3 > gpurun_out/r2_sweep_apply$v.jsonl 2>gpurun_out/r2_sweep_apply$v.err; echo "variant '$v'"; cut -c1-330 gpurun_out/r2_sweep_apply$v.jsonl
done
timeout 900 python bench.py --workload config3 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-files > gpurun_out/r2_bench_c3_v8.log 2> gpurun_out/r2_bench_c3_v8.err; tail -1 gpurun_out/r2_bench_c3_v8.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print(d['value'], d['ms_per_step'], r['kernel'], r['kernel_ms'], r['apply'], r['frac'], r['frac_step'], d['checks'])"
