set -x
cd /root/repo
mkdir -p gpurun_out
export KB_WATCHDOG_S=30
timeout 900 python -m pytest tests/test_evict_parity.py -x -q -m gpu 2>&1 | tail -15
timeout 600 python tools/cycle_time.py c3 0.3 1 > gpurun_out/r02_cycle_c3.json 2> gpurun_out/r02_cycle_c3.err; tail -c 1500 gpurun_out/r02_cycle_c3.json; tail -3 gpurun_out/r02_cycle_c3.err
