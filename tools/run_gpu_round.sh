set -x
TAG=${1:-b}
timeout 900 python -m pytest tests -m gpu -q -s -rs -x 2>&1 | grep -E "status|plane init|cov rel err|sharded|passed|failed|SKIP|FAILED|Error|error|assert" | cut -c1-250
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_$TAG.json 2> gpurun_out/bench_r1_$TAG.err; tail -c 300 gpurun_out/bench_r1_$TAG.json; tail -5 gpurun_out/bench_r1_$TAG.err
