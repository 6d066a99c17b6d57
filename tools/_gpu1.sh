set -x
cd /root/repo
timeout 120 python tools/quick_time.py c2 2 2>&1 | tail -12
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "reference_allocate or baseline_configs or repeatable" 2>&1 | tail -15
timeout 200 python tools/quick_time.py c3 3 2>&1 | tail -14
timeout 200 python tools/quick_time.py c3 2 32 2>&1 | tail -8
