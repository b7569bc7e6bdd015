for rep in 1 2; do
for v in base ordsimple; do
L=$PWD/kmer_mapper_b200/libkmer_mapper_b200.so; [ $v = ordsimple ] && L=$PWD/kmer_mapper_b200/libkmer_mapper_b200_ordsimple.so
KMB_LIB_PATH=$L timeout 600 python tools/sweep.py --workload config3 --reads 25000000 --grid r2 --steps 3 2> gpurun_out/r2_c3_$v.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$v', 'kernel_ms', round(d['kernel_ms'],3), 'step_ms', round(d['step_ms'],3), d['cand_per_kmer'], d['counts_equal_first'])"
done
done
