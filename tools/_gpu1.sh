cd /root/repo
export KB_WATCHDOG_S=120
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/r02f_bench.err | tail -1 > gpurun_out/r02f_bench_n1.json
cut -c1-400 gpurun_out/r02f_bench_n1.json
KB_PIPE_TIMING=1 timeout 100 python tools/quick_time.py c3 2 2>&1 | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02f_cycle_launches.csv python tools/quick_time.py c3 2 > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:cycle_kernel -c 1 -o gpurun_out/r02f_cycle -f python tools/quick_time.py c3 1 > gpurun_out/r02f_ncu.log 2>&1
ls -la gpurun_out/r02f_cycle.ncu-rep
timeout 600 python tools/c5_properties.py c5 > gpurun_out/r02f_c5_properties_n1.json 2>gpurun_out/r02f_c5.err; tail -c 600 gpurun_out/r02f_c5_properties_n1.json
timeout 600 python tools/c5_properties.py c4 > gpurun_out/r02f_c4_properties_n1.json 2>gpurun_out/r02f_c4.err; tail -c 400 gpurun_out/r02f_c4_properties_n1.json
timeout 600 python tools/cycle_time.py c3 0.3 0 > gpurun_out/r02f_cycle_c3_n1.json 2>gpurun_out/r02f_cyc.err; tail -c 400 gpurun_out/r02f_cycle_c3_n1.json
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
