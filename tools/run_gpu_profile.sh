set -x
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_final2.csv python tools/profile_step.py > gpurun_out/prof_launch.log 2>&1; tail -1 gpurun_out/prof_launch.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:chol_fused_kernel -s 16 -c 2 -o gpurun_out/prof_r1_chol2 -f python tools/profile_step.py > gpurun_out/prof_chol.log 2>&1; tail -2 gpurun_out/prof_chol.log
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:"gram_kernel|feature_kernel|finish_update_kernel" -s 21 -c 3 -o gpurun_out/prof_r1_point2 -f python tools/profile_step.py > gpurun_out/prof_point.log 2>&1; tail -2 gpurun_out/prof_point.log
timeout 100 python tools/microbench_chol.py 474 512 2>&1 | grep "fused chol\|phase" | head -2 | cut -c1-1000
timeout 600 python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; tail -c 300 gpurun_out/bench_r1_final.json
