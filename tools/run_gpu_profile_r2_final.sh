# Round-2 final evidence run (one GPU): launch list under ncu, --set full captures of the dominant kernels and of the plane-fitting kernels,
# compute-sanitizer on the new kernels, the default bench and the reference arm.  Outputs land in gpurun_out/ (summaries are copied into
# profiles/ by hand / tools/summarize_launches.py).
set -x
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches.csv python tools/profile_step.py > gpurun_out/r2f_prof_launch.log 2>&1; tail -1 gpurun_out/r2f_prof_launch.log
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches_planefit.csv python tools/profile_planefit.py > gpurun_out/r2f_prof_launch_pf.log 2>&1; tail -1 gpurun_out/r2f_prof_launch_pf.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:chol_fused_kernel -s 16 -c 2 -o gpurun_out/r2f_chol -f python tools/profile_step.py > gpurun_out/r2f_prof_chol.log 2>&1; tail -1 gpurun_out/r2f_prof_chol.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"msckf_feature_warp_kernel|gram_kernel|gram_dpart_kernel|gemm_f64_kernel" -s 40 -c 8 -o gpurun_out/r2f_point -f python tools/profile_step.py > gpurun_out/r2f_prof_point.log 2>&1; tail -1 gpurun_out/r2f_prof_point.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"plane_ransac_kernel|plane_ransac_select_kernel|optimize_plane_kernel" -c 3 -o gpurun_out/r2f_planefit -f python tools/profile_planefit.py > gpurun_out/r2f_prof_planefit.log 2>&1; tail -1 gpurun_out/r2f_prof_planefit.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_planefit.py tests/test_gpu_anchors.py -q -x -k "small_planes or tiny or rep" > gpurun_out/r2f_sanitizer.log 2>&1; echo "sanitizer rc $?"; tail -4 gpurun_out/r2f_sanitizer.log
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 300 gpurun_out/r2f_bench.json
timeout 900 python bench.py --impl reference > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; tail -c 400 gpurun_out/r2f_bench_ref.json
