timeout 600 python tools/sweep.py --workload config3 --reads 25000000 --grid r2 --steps 3 2> gpurun_out/r2_c3_cold.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('cold', 'kernel_ms', round(d['kernel_ms'],3), 'step_ms', round(d['step_ms'],3), d['cand_per_kmer'], d['counts_equal_first'])"
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_t29.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t29.log); tail -2 gpurun_out/r2_t29.log
