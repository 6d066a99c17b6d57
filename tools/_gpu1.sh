set -x
cd /root/repo
mkdir -p gpurun_out
export KB_WATCHDOG_S=30
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "preferred or unsupported or baseline or reference_allocate" 2>&1 | tail -15
timeout 100 python tools/quick_time.py c3 2 2>&1 | tail -6
