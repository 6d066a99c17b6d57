set -x
cd /root/repo
mkdir -p gpurun_out
export KB_WATCHDOG_S=60
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err; tail -c 300 gpurun_out/r02e_bench_n1.json; tail -3 gpurun_out/r02e_bench_n1.err
timeout 900 python tools/cycle_time.py c3 0.3 1 > gpurun_out/r02e_cycle_c3.json 2> gpurun_out/r02e_cycle_c3.err; python -c "
import json;d=json.load(open('gpurun_out/r02e_cycle_c3.json'));print({k:v for k,v in d['rep1'].items() if 'bounds' not in k}); print(d['oracle'])"; tail -3 gpurun_out/r02e_cycle_c3.err
export KB_WATCHDOG_S=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02e_launches.csv python tools/quick_time.py c3 2 > gpurun_out/ncu_l.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:cycle_kernel -c 1 -o gpurun_out/r02e_cycle python tools/quick_time.py c3 1 > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out/ | tail -5
