cd /root/repo
export KB_WATCHDOG_S=120
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_evict_parity.py tests/test_gpu_affinity.py -m gpu -q -x -k "not big_affinity and not baseline_size and not larger" 2>&1 | tail -3
