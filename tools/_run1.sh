set -x
timeout 300 python tools/debug_nbits32.py 32 2>&1 | cut -c1-1700 | tail -30
QOB_QREG_CHAIN_BITS=0 timeout 300 python tools/debug_nbits32.py 32 2>&1 | grep "MATCH\|rank" | cut -c1-300
timeout 300 python bench.py --spins 32 --steps 2 --warmup 3 --no-extra-configs 2> gpurun_out/n32.err | cut -c1-1200; grep "\[bench\]" gpurun_out/n32.err
