(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_t33.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t33.log); tail -2 gpurun_out/r2_t33.log
(KMB_LIB_PATH=$PWD/kmer_mapper_b200/libkmer_mapper_b200_bounds.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_t33b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t33b.log); tail -2 gpurun_out/r2_t33b.log
for rep in 1 2; do
for v in nobulk bulk; do
L=$PWD/kmer_mapper_b200/libkmer_mapper_b200.so; [ $v = nobulk ] && L=$PWD/kmer_mapper_b200/libkmer_mapper_b200_nobulk.so
KMB_LIB_PATH=$L timeout 600 python bench.py --no-files --no-e2e --no-cpu-baseline --no-oracle --steps 4 --warmup 2 2> gpurun_out/r2_bulk_$v.err | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('config2 $v', round(d['value'],2), 'step', round(d['ms_per_step'],3), 'kernel', round(r['kernel_ms'],3), 'apply', round(r['apply']['ms'],3))"
done
done
for w in config5 config4_k21; do
for v in nobulk bulk; do
L=$PWD/kmer_mapper_b200/libkmer_mapper_b200.so; [ $v = nobulk ] && L=$PWD/kmer_mapper_b200/libkmer_mapper_b200_nobulk.so
KMB_LIB_PATH=$L timeout 600 python bench.py --workload $w --no-files --no-e2e --no-cpu-baseline --no-oracle --steps 3 --warmup 2 2>> gpurun_out/r2_bulk_$v.err | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$w $v', round(d['value'],2), 'step', round(d['ms_per_step'],3), 'kernel', round(r['kernel_ms'],3), 'apply', round(r['apply']['ms'],3))"
done
done
