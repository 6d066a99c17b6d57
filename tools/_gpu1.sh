set -x
cd /root/repo
mkdir -p gpurun_out
export KB_WATCHDOG_S=30
KB_PIPE_TIMING=1 timeout 100 python tools/quick_time.py c3 2 2>&1 | tail -4
timeout 100 python tools/quick_time.py c3 3 2>&1 | grep rep
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
