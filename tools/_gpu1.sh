set -x
cd /root/repo
mkdir -p gpurun_out
export KB_WATCHDOG_S=10
KB_PIPE_DEBUG=1 timeout 60 python tools/quick_time.py c2 2 2>&1 | tail -12
timeout 100 python tools/quick_time.py c3 3 2>&1 | tail -16
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15
