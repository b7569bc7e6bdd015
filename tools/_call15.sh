run() { KMB_LIB_PATH=$PWD/kmer_mapper_b200/libkmer_mapper_b200$1.so timeout 400 python tools/sweep.py --grid $2 --steps 3 > gpurun_out/r2_sweep_rec$1_$2.jsonl 2>gpurun_out/r2_sweep_rec$1_$2.err; echo "variant '$1' grid $2"; cut -c1-330 gpurun_out/r2_sweep_rec$1_$2.jsonl; tail -2 gpurun_out/r2_sweep_rec$1_$2.err; }
run _rec u
run _recs64 u
