cd /root/repo
export KB_WATCHDOG_S=30
KB_PIPE_TIMING=1 timeout 100 python tools/quick_time.py c3 2 2>&1 | tail -4
timeout 100 python tools/quick_time.py c3 4 2>&1 | grep -E "rep[123]" | cut -c1-70
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py 2>&1 | tail -2 | cut -c1-1500
