cd /root/repo
export KB_WATCHDOG_S=120
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_affinity.py tests/test_host_cpp.py -m gpu -q -x 2>&1 | tail -4
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "matrix or score or best_nodes or fail_loudly" 2>&1 | tail -2
