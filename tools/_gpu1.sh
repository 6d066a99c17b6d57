cd /root/repo
export KB_WATCHDOG_S=120
mkdir -p gpurun_out
timeout 215 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
