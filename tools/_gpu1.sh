set -x
cd /root/repo
mkdir -p gpurun_out
export KB_WATCHDOG_S=30
KB_PIPE_TIMING=1 timeout 100 python tools/quick_time.py c3 2 2>&1 | tail -4
timeout 100 python tools/quick_time.py c3 3 2>&1 | grep rep
timeout 900 python -m pytest tests/test_evict_parity.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python tools/cycle_time.py c3 0.3 0 > gpurun_out/r02b_cycle_c3.json 2> gpurun_out/r02b_cycle_c3.err; python -c "
import json;d=json.load(open('gpurun_out/r02b_cycle_c3.json'));print(d['workload']);print({k:v for k,v in d['rep1'].items() if 'bounds' not in k})"; tail -3 gpurun_out/r02b_cycle_c3.err
timeout 900 python tools/cycle_time.py c4 0.3 0 > gpurun_out/r02b_cycle_c4.json 2> gpurun_out/r02b_cycle_c4.err; python -c "
import json;d=json.load(open('gpurun_out/r02b_cycle_c4.json'));print(d['workload']);print(d['load_ms']);print({k:v for k,v in d['rep1'].items() if 'bounds' not in k})"; tail -3 gpurun_out/r02b_cycle_c4.err
