mkdir -p gpurun_out/r3i
N=$(nvidia-smi -L | wc -l)
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551"
(time timeout 600 $T bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r3i/bench_n${N}_cfg3.json 2> gpurun_out/r3i/bench_n${N}_cfg3.err) 2>&1 | grep real
python - <<PY
import json
d=json.loads(open('gpurun_out/r3i/bench_n${N}_cfg3.json').read().strip().splitlines()[-1])
print('N=${N}', d['value'], d['ms_per_step'], d['e2e']['value'], d['resident_device_only']['value'], d['verified'], d['clocks'])
PY
