mkdir -p gpurun_out
( timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pairs.py tests/test_gpu_ingest_eval.py -q -x -k "mixed_shapes or second_fast_pass or svdpp_warp or pairs_have or side_features_exact or basic_k64 or general_k40 or svdpp_k64_tags or device_eval_ugroup or conflict_free_is_exact" 2>&1 | tail -15 ) > gpurun_out/sanitizer.log 2>&1
echo "sanitizer rc=$?"; tail -12 gpurun_out/sanitizer.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mf -s 5 -c 1 -o gpurun_out/kmf_ni2_c4 python tools/bench_configs.py c4 --scale 0.3 > gpurun_out/ncu_c4b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_svdpp -s 2 -c 1 -o gpurun_out/ksvdpp_c3 python tools/bench_configs.py c3 --scale 0.1 > gpurun_out/ncu_c3c.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
