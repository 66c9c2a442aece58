set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
export QOB_BENCH_SKIP_E2E=1 QOB_DIST_EXCHANGE=capi QOB_DIST_TRACE=1
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --spins 32 --steps 2 --warmup 3 --no-extra-configs > gpurun_out/t1.json 2> gpurun_out/t1.err; grep "trace\|\[bench\]" gpurun_out/t1.err | tail -6
QOB_DIST_DIRECT_SPLIT=1 timeout 300 $TR --master-port 29514 bench.py --gpus 2 --spins 32 --steps 2 --warmup 3 --no-extra-configs > gpurun_out/t2.json 2> gpurun_out/t2.err; grep "trace\|\[bench\]" gpurun_out/t2.err | tail -6
QOB_DIST_DIRECT_SPLIT=1 timeout 300 $TR --master-port 29515 bench.py --gpus 2 --spins 33 --steps 2 --warmup 3 --no-extra-configs > gpurun_out/t3.json 2> gpurun_out/t3.err; grep "trace\|\[bench\]" gpurun_out/t3.err | tail -6
