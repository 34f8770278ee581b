SFMM_TRACE_HOST=1 timeout 120 python tools/e2e_breakdown.py binary 50 5000 2>&1 | tail -4
SFMM_TRACE_HOST=1 timeout 120 python tools/e2e_breakdown.py float 60 8000 2>&1 | tail -4
