mkdir -p gpurun_out/r2v
timeout 900 python -m pytest tests/test_parity_binary.py -m gpu -x -q 2>&1 | tail -5
B="--steps 5 --warmup 3 --no-extra --no-alt-engine --no-cpu-baseline --verify 2 --device-only-iters 3"
for w in cfg2 cfg3; do
SFMM_NO_F4=1 timeout 300 python bench.py --workload $w $B > gpurun_out/r2v/${w}_i8.json 2> gpurun_out/r2v/${w}_i8.err
timeout 300 python bench.py --workload $w $B > gpurun_out/r2v/${w}_f4.json 2> gpurun_out/r2v/${w}_f4.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2v/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], d['value'], d.get('resident_device_only',{}).get('value'), d['roofline'].get('frac'), d.get('verified'))
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/r2v/cfg2_f4.err
