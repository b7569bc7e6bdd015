for v in base b12s35 b12s40 base b12s35; do
L=$PWD/kmer_mapper_b200/libkmer_mapper_b200.so; [ $v != base ] && L=$PWD/kmer_mapper_b200/libkmer_mapper_b200_$v.so
KMB_LIB_PATH=$L timeout 600 python tools/sweep.py --workload config3 --reads 50000000 --grid r2 --steps 3 2> gpurun_out/r2_c3_$v.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$v', 'kernel_ms', round(d['kernel_ms'],3), 'apply_ms', round(d['apply_ms'],3), 'step_ms', round(d['step_ms'],3), d['cand_per_kmer'], d['counts_equal_first'])"
done
(KMB_LIB_PATH=$PWD/kmer_mapper_b200/libkmer_mapper_b200_b12s35.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_t37.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t37.log); tail -2 gpurun_out/r2_t37.log
