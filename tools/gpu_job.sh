mkdir -p gpurun_out/r3d
timeout 600 python -m pytest tests/test_orb_gpu.py tests/test_cpp_adapter.py -m gpu -x -q 2>&1 | tail -3
timeout 120 python tools/orb_time.py
SFMM_ORB_NO_GRAPH=1 timeout 120 python tools/orb_time.py
SFMM_ORB_NO_GRAPH=1 timeout 600 python -m pytest tests/test_orb_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r3d/launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --verify 0 --device-only-iters 1 --e2e-steps 1 --e2e-warmup 1 > gpurun_out/r3d/launches_bench.log 2>&1
tail -1 gpurun_out/r3d/launches_bench.log | cut -c1-300
wc -l gpurun_out/r3d/launches_default.csv
