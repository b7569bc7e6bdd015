set -x
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t11.log); tail -5 gpurun_out/r2_t11.log
timeout 900 python bench.py > gpurun_out/r2_bench_c2_hybrid.log 2> gpurun_out/r2_bench_c2_hybrid.err; tail -c 6000 gpurun_out/r2_bench_c2_hybrid.log; tail -3 gpurun_out/r2_bench_c2_hybrid.err
for hp in 1 0; do timeout 600 python bench.py --no-files --no-cpu-baseline --no-oracle --host-pack $hp > gpurun_out/r2_bench_c2_hp$hp.log 2>&1; python - <<PY
import json
l=[x for x in open("gpurun_out/r2_bench_c2_hp$hp.log") if x.startswith("{")][-1]
print("host_pack=$hp", json.dumps(json.loads(l)["e2e"]))
PY
done
