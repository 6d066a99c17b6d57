cd /root/repo
export KB_WATCHDOG_S=120
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_affinity.py -m gpu -q -x -k "shipped_action_list or refusals or pipeline" 2>&1 | tail -3
timeout 300 python -m pytest tests/test_evict_parity.py -m gpu -q -x -k "action_lists or reference_action" 2>&1 | tail -2
