mkdir -p gpurun_out/r3e
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
B="--steps 5 --warmup 3 --no-extra --no-alt-engine --no-cpu-baseline --verify 2 --device-only-iters 3"
for w in temple_akaze temple_sift; do
SFMM_BENCH_TRACE=1 timeout 200 python bench.py --workload $w $B > gpurun_out/r3e/$w.json 2> gpurun_out/r3e/$w.err; grep "e2e\]" gpurun_out/r3e/$w.err | tail -2
python -c "import json; d=json.loads(open('gpurun_out/r3e/$w.json').read().strip().splitlines()[-1]); print('$w', d['value'], d['e2e']['value'], d['resident_device_only']['value'], d['ms_per_step'])"
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r3e/launches_default.csv python bench.py --steps 2 --warmup 1 --no-extra --no-alt-engine --no-cpu-baseline --verify 0 --device-only-iters 1 --e2e-steps 1 --e2e-warmup 1 > gpurun_out/r3e/launches_bench.log 2>&1
wc -l gpurun_out/r3e/launches_default.csv
