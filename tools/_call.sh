set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551"
timeout 300 $TR tools/sharded_check.py small > gpurun_out/r02h_sharded_small.log 2>&1; echo rc=$? >> gpurun_out/r02h_sharded_small.log
timeout 300 $TR tools/sharded_check.py cfg2 > gpurun_out/r02h_sharded_cfg2.log 2>&1; echo rc=$? >> gpurun_out/r02h_sharded_cfg2.log
timeout 300 $TR tools/sharded_check.py cfg4 > gpurun_out/r02h_sharded_cfg4.log 2>&1; echo rc=$? >> gpurun_out/r02h_sharded_cfg4.log
timeout 300 $TR tools/sharded_profile.py cfg4 > gpurun_out/r02h_sharded_timeline_2gpu.log 2>&1
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02h_bench_2gpu.json 2> gpurun_out/r02h_bench_2gpu.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02h_bench_1gpu.json 2> gpurun_out/r02h_bench_1gpu.err
tail -n 3 gpurun_out/r02h_sharded_*.log
