# Round-2 evidence run (one GPU): launch list under ncu, --set full captures of the three dominant kernels, the default bench and the
# reference arm.  Outputs land in gpurun_out/ (summaries are copied into profiles/ by hand / tools/summarize_launches.py).
set -x
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python tools/profile_step.py > gpurun_out/r2_prof_launch.log 2>&1; tail -1 gpurun_out/r2_prof_launch.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:chol_fused_kernel -s 16 -c 2 -o gpurun_out/r2_chol_final -f python tools/profile_step.py > gpurun_out/r2_prof_chol.log 2>&1; tail -1 gpurun_out/r2_prof_chol.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"msckf_feature_warp_kernel|gram_kernel|gram_dpart_kernel|gemm_f64_kernel" -s 40 -c 8 -o gpurun_out/r2_point_final -f python tools/profile_step.py > gpurun_out/r2_prof_point.log 2>&1; tail -1 gpurun_out/r2_prof_point.log
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -c 400 gpurun_out/r2_bench_final.json
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; tail -c 600 gpurun_out/r2_bench_ref.json
