B="--steps 1 --warmup 1 --no-extra --no-alt-engine --no-cpu-baseline --verify 2 --device-only-iters 1 --e2e-steps 3 --e2e-warmup 1"
SFMM_BENCH_TRACE=1 SFMM_TRACE_HOST=1 timeout 200 python bench.py $B 2>&1 | grep "e2e\]\|sfmm\]\|verified" | tail -5 | cut -c1-250
SFMM_BENCH_TRACE=1 timeout 200 python bench.py --workload cfg4s $B 2>&1 | grep "e2e\]" | tail -2
SFMM_BENCH_TRACE=1 timeout 200 python bench.py --workload cfg2 $B 2>&1 | grep "e2e\]" | tail -2
timeout 600 python -m pytest tests/test_parity_binary.py tests/test_next_rows.py tests/test_cpp_adapter.py -m gpu -x -q 2>&1 | tail -2
