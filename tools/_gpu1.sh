cd /root/repo
export KB_WATCHDOG_S=120
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_affinity.py -x -q 2>&1 | tail -15
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "baseline or random_sessions or preferred or bind_list" 2>&1 | tail -4
timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 2 2>gpurun_out/r02g_bench.err | tail -1 > gpurun_out/r02g_bench_n1.json
cut -c1-300 gpurun_out/r02g_bench_n1.json
