cd /root/repo
export KB_WATCHDOG_S=120
mkdir -p gpurun_out
P=r02l
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/${P}_bench.err | tail -1 > gpurun_out/${P}_bench_n1.json
cut -c1-260 gpurun_out/${P}_bench_n1.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
