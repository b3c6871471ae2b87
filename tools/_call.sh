set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551"
DCL_BENCH_DUMP_STEPS=1 timeout 600 $TR bench.py --gpus 2 --steps 60 --warmup 5 > gpurun_out/r02p_a.json 2> gpurun_out/r02p_a.err
NCCL_PROTO=Simple DCL_BENCH_DUMP_STEPS=1 timeout 600 $TR bench.py --gpus 2 --steps 60 --warmup 5 > gpurun_out/r02p_b.json 2> gpurun_out/r02p_b.err
NCCL_PROTO=Simple timeout 300 $TR tools/sharded_profile.py cfg4 > gpurun_out/r02p_sharded_timeline_2gpu_simple.log 2>&1
grep "per-step" gpurun_out/r02p_*.err
