set -x
cd /root/repo
mkdir -p gpurun_out
export KB_WATCHDOG_S=60
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02d_bench_n1.json 2> gpurun_out/r02d_bench_n1.err; tail -c 1200 gpurun_out/r02d_bench_n1.json; tail -3 gpurun_out/r02d_bench_n1.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02d_bench_reference_n1.json 2> gpurun_out/r02d_ref.err; tail -c 900 gpurun_out/r02d_bench_reference_n1.json
export KB_WATCHDOG_S=0
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/quick_time.py c2 1 > gpurun_out/r02d_memcheck_pipeline_c2.log 2>&1; tail -5 gpurun_out/r02d_memcheck_pipeline_c2.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/quick_time.py c1 1 > gpurun_out/r02d_racecheck_pipeline_c1.log 2>&1; tail -5 gpurun_out/r02d_racecheck_pipeline_c1.log
KB_PIPE=0 timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/quick_time.py c2 1 > gpurun_out/r02d_racecheck_visit_c2.log 2>&1; tail -5 gpurun_out/r02d_racecheck_visit_c2.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_evict_parity.py -q -m gpu -k "reference_action or synthetic" > gpurun_out/r02d_memcheck_evict.log 2>&1; tail -5 gpurun_out/r02d_memcheck_evict.log
