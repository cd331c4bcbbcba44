set -x
timeout 600 python -m pytest tests -m gpu -q -rs -x 2>&1 | tail -4 | cut -c1-200
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_final.csv python tools/profile_step.py > gpurun_out/prof_launch.log 2>&1; tail -2 gpurun_out/prof_launch.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:chol_fused_kernel -s 16 -c 2 -o gpurun_out/prof_r1_chol -f python tools/profile_step.py > gpurun_out/prof_chol.log 2>&1; tail -3 gpurun_out/prof_chol.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gram_kernel|feature_kernel" -s 16 -c 2 -o gpurun_out/prof_r1_point -f python tools/profile_step.py > gpurun_out/prof_point.log 2>&1; tail -3 gpurun_out/prof_point.log
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:gemm_f64_kernel -s 24 -c 3 -o gpurun_out/prof_r1_gemm -f python tools/profile_step.py > gpurun_out/prof_gemm.log 2>&1; tail -3 gpurun_out/prof_gemm.log
ls -la gpurun_out/*.ncu-rep
