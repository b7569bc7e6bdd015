for rep in 1 2; do
for v in cold inl; do
L=$PWD/kmer_mapper_b200/libkmer_mapper_b200.so; [ $v = inl ] && L=$PWD/kmer_mapper_b200/libkmer_mapper_b200_inl.so
KMB_LIB_PATH=$L timeout 600 python tools/sweep.py --workload config3 --reads 25000000 --grid r2 --steps 3 2> gpurun_out/r2_c3_$v.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$v', 'kernel_ms', round(d['kernel_ms'],3), 'step_ms', round(d['step_ms'],3), d['cand_per_kmer'], d['counts_equal_first'])"
done
done
for s in 0 1 0 1; do
timeout 600 python bench.py --no-files --no-cpu-baseline --no-oracle --steps 4 --warmup 2 --opt host_pack_streaming=$s 2> gpurun_out/r2_e2e_nt$s.err | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']; print('streaming=$s', 'e2e', round(e['value'],2), 'ms', round(e['ms_per_step'],2), e['host_transport'][:40], 'host_read', e.get('host_read_GBps_per_rank_all_ranks_at_once'), 'value', round(d['value'],1))"
done
for s in 0 1; do
timeout 600 python bench.py --no-files --no-cpu-baseline --no-oracle --steps 4 --warmup 2 --host-pack 1 --opt host_pack_streaming=$s 2> gpurun_out/r2_e2e_hp1_nt$s.err | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']; print('packed-only streaming=$s', 'e2e', round(e['value'],2), 'ms', round(e['ms_per_step'],2))"
done
