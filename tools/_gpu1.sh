set -x
cd /root/repo
mkdir -p gpurun_out
export KB_WATCHDOG_S=30
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
timeout 600 python tools/cycle_time.py c3 0.3 0 > gpurun_out/r02c_cycle_c3.json 2> gpurun_out/r02c_cycle_c3.err; python -c "
import json;d=json.load(open('gpurun_out/r02c_cycle_c3.json'));print(d['workload']);print({k:v for k,v in d['rep1'].items() if 'bounds' not in k})"; tail -3 gpurun_out/r02c_cycle_c3.err
timeout 900 python tools/cycle_time.py c4 0.3 0 > gpurun_out/r02c_cycle_c4.json 2> gpurun_out/r02c_cycle_c4.err; python -c "
import json;d=json.load(open('gpurun_out/r02c_cycle_c4.json'));print(d['workload']);print(d['load_ms']);print({k:v for k,v in d['rep1'].items() if 'bounds' not in k})"; tail -3 gpurun_out/r02c_cycle_c4.err
