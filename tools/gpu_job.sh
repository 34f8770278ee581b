SFMM_CHUNKS=16 timeout 120 python tools/repro_tmp.py 30 10000 2>&1 | tail -1 | cut -c1-200
timeout 120 python tools/repro_tmp.py 130 10000 2>&1 | tail -1 | cut -c1-200
timeout 600 compute-sanitizer --tool memcheck python tools/repro_tmp.py 130 10000 2>&1 | grep -v "^$" | head -30 | cut -c1-300
