for w in config4_k15 config4_k21 config2; do
for v in casfirst new; do
L=$PWD/kmer_mapper_b200/libkmer_mapper_b200.so; [ $v = casfirst ] && L=$PWD/kmer_mapper_b200/libkmer_mapper_b200_casfirst.so
KMB_LIB_PATH=$L timeout 600 python tools/sweep.py --workload $w --reads 25000000 --grid slabs --steps 3 2> gpurun_out/r2_apply_$v.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$w $v', d['opts'], 'apply_ms', round(d['apply_ms'],3), 'kernel_ms', round(d['kernel_ms'],3), 'step_ms', round(d['step_ms'],3), d['counts_equal_first'])"
done
done
