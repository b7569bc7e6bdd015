# usage: bash tools/_callN.sh N
N=$1
P=29611
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_v8_c2_n$N.log 2> gpurun_out/r2_v8_c2_n$N.err; tail -1 gpurun_out/r2_v8_c2_n$N.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('c2 N=$N', d['value'], d['ms_per_step'], json.dumps(d['e2e']), d['checks'])"; tail -2 gpurun_out/r2_v8_c2_n$N.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus $N --workload config3 --steps 3 --warmup 2 --no-e2e > gpurun_out/r2_v8_c3_n$N.log 2> gpurun_out/r2_v8_c3_n$N.err; tail -1 gpurun_out/r2_v8_c3_n$N.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('c3 N=$N', d['value'], d['ms_per_step'], d['roofline']['frac_step'], d['checks'])"; tail -2 gpurun_out/r2_v8_c3_n$N.err
nvidia-smi topo -m > gpurun_out/r2_topo_n$N.txt 2>&1; lscpu | head -25 >> gpurun_out/r2_topo_n$N.txt
