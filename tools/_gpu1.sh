set -x
cd /root/repo
export KB_WATCHDOG_S=30
KB_PIPE_TIMING=1 timeout 100 python tools/quick_time.py c3 2 2>&1 | tail -6
