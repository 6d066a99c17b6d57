cd /root/repo
export KB_WATCHDOG_S=120
mkdir -p gpurun_out
P=r02i
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --cpu-seconds 2 2>gpurun_out/${P}_bench_n2.err | tail -1 > gpurun_out/${P}_bench_n2.json
cut -c1-420 gpurun_out/${P}_bench_n2.json; tail -3 gpurun_out/${P}_bench_n2.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --cpu-seconds 1 --sessions replicated 2>gpurun_out/${P}_bench_n2_repl.err | tail -1 > gpurun_out/${P}_bench_n2_replicated.json
cut -c1-420 gpurun_out/${P}_bench_n2_replicated.json
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
