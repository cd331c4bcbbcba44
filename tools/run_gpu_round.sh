set -x
TAG=${1:-b}
python -m pytest tests -m gpu -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_$TAG.json 2> gpurun_out/bench_r1_$TAG.err; tail -c 2300 gpurun_out/bench_r1_$TAG.json; tail -5 gpurun_out/bench_r1_$TAG.err
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python tools/profile_step.py > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log; wc -l gpurun_out/launches_$TAG.csv
