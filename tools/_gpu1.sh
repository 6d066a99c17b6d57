cd /root/repo
export KB_WATCHDOG_S=120
mkdir -p gpurun_out
P=r02j
timeout 600 python -m pytest tests/test_gpu_affinity.py -q -x -s 2>&1 | grep -E "passed|failed|timing|Error" | tail -8
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "backfill or phantom" 2>&1 | tail -2
timeout 200 python tools/affinity_run.py > gpurun_out/${P}_affinity_run.json 2>gpurun_out/${P}_aff.err; cut -c1-330 gpurun_out/${P}_affinity_run.json
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/affinity_run.py 300 120 1 > gpurun_out/${P}_memcheck_affinity.txt 2>&1; tail -4 gpurun_out/${P}_memcheck_affinity.txt
timeout 400 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python tools/affinity_run.py 300 120 1 > gpurun_out/${P}_racecheck_affinity.txt 2>&1; tail -4 gpurun_out/${P}_racecheck_affinity.txt
