mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ugroup -s 2 -c 1 -o gpurun_out/kugroup_c3b python tools/bench_configs.py c3 --scale 0.2 > gpurun_out/ncu_c3.log 2>&1
tail -2 gpurun_out/ncu_c3.log
