set -x
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q -s) > gpurun_out/r2f_tests.log 2>&1; tail -4 gpurun_out/r2f_tests.log | cut -c1-200
bash tools/run_gpu_profile_r2_final.sh 2>&1 | tail -25
