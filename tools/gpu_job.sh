mkdir -p gpurun_out/r2n
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2n/launches_default.csv python bench.py --steps 2 --warmup 1 --no-extra --no-alt-engine --no-cpu-baseline --verify 0 --device-only-iters 1 --e2e-steps 1 --e2e-warmup 1 > gpurun_out/r2n/launches_bench.log 2>&1
tail -2 gpurun_out/r2n/launches_bench.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tensor_knn2_ts -s 2 -c 1 -o gpurun_out/r2n/ts_i8p_cfg3 python bench.py --steps 1 --warmup 1 --no-extra --no-alt-engine --no-cpu-baseline --verify 0 --device-only-iters 1 --e2e-steps 1 --e2e-warmup 0 > gpurun_out/r2n/ncu1.log 2>&1; tail -2 gpurun_out/r2n/ncu1.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tensor_knn2_ts -s 2 -c 1 -o gpurun_out/r2n/ts_f16x_cfg4s python bench.py --workload cfg4s --steps 1 --warmup 1 --no-extra --no-alt-engine --no-cpu-baseline --verify 0 --device-only-iters 1 --e2e-steps 1 --e2e-warmup 0 > gpurun_out/r2n/ncu2.log 2>&1; tail -2 gpurun_out/r2n/ncu2.log | cut -c1-200
SFMM_BENCH_TRACE=1 timeout 100 python bench.py --steps 3 --warmup 1 --no-extra --no-alt-engine --no-cpu-baseline --verify 0 --device-only-iters 1 2>&1 | grep "e2e\]" | tail -2
ls -la gpurun_out/r2n
