set -x
TAG=${1:-b}
timeout 120 python tools/microbench_chol.py 2>&1 | grep -vE "^  [DP][0-9]" 
timeout 600 python -m pytest tests -m gpu -q -rs -x 2>&1 | tail -6 | cut -c1-250
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_$TAG.json 2> gpurun_out/bench_r1_$TAG.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r1_$TAG.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
print(d["roofline"]["per_kernel_ms_per_step"], d["roofline"]["per_kernel_launches_per_step"])
print(d.get("concurrent_filters"))
PY
tail -3 gpurun_out/bench_r1_$TAG.err
