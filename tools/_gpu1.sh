cd /root/repo
export KB_WATCHDOG_S=120
mkdir -p gpurun_out
P=r02h
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/${P}_bench.err | tail -1 > gpurun_out/${P}_bench_n1.json
cut -c1-300 gpurun_out/${P}_bench_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/${P}_bench_reference_n1.json
cut -c1-200 gpurun_out/${P}_bench_reference_n1.json
timeout 200 python tools/affinity_run.py > gpurun_out/${P}_affinity_run.json 2>gpurun_out/${P}_aff.err; cat gpurun_out/${P}_affinity_run.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${P}_cycle_launches.csv python tools/quick_time.py c3 2 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${P}_affinity_launches.csv python tools/affinity_run.py 2000 1000 1 > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:cycle_kernel -c 1 -o gpurun_out/${P}_cycle -f python tools/quick_time.py c3 1 > gpurun_out/${P}_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:visit_kernel -s 50 -c 1 -o gpurun_out/${P}_visit_aff -f python tools/affinity_run.py 2000 1000 1 > gpurun_out/${P}_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 400 python tools/c5_properties.py c5 > gpurun_out/${P}_c5_properties_n1.json 2>gpurun_out/${P}_c5.err; tail -c 400 gpurun_out/${P}_c5_properties_n1.json
timeout 300 python tools/c5_properties.py c4 > gpurun_out/${P}_c4_properties_n1.json 2>gpurun_out/${P}_c4.err; tail -c 300 gpurun_out/${P}_c4_properties_n1.json
