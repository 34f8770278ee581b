mkdir -p gpurun_out/r3c
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
(timeout 500 compute-sanitizer --tool memcheck python tools/sanitizer_workload.py > gpurun_out/r3c/memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r3c/memcheck.log); tail -3 gpurun_out/r3c/memcheck.log
(timeout 700 compute-sanitizer --tool racecheck python tools/sanitizer_workload.py > gpurun_out/r3c/racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/r3c/racecheck.log); tail -3 gpurun_out/r3c/racecheck.log
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
