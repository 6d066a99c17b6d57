cd /root/repo
export KB_WATCHDOG_S=120
mkdir -p gpurun_out
P=r02i
timeout 600 python -m pytest tests/test_gpu_affinity.py -q -x 2>&1 | tail -3
timeout 200 python tools/affinity_run.py > gpurun_out/${P}_affinity_run.json 2>gpurun_out/${P}_aff.err; cat gpurun_out/${P}_affinity_run.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${P}_affinity_launches.csv python tools/affinity_run.py 2000 1000 1 > /dev/null 2>&1
