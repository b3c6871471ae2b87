set -x
DCL_FILL_PLACEMENT=1 python tools/torch_profile.py cfg2 > gpurun_out/r02y_timeline_cfg2_late.log 2>&1
DCL_FILL_PLACEMENT=1 python tools/torch_profile.py cfg4 > gpurun_out/r02y_timeline_cfg4_late.log 2>&1
DCL_FILL_PLACEMENT=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-hbm > gpurun_out/r02y_bench_late.json 2> gpurun_out/r02y_bench_late.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-hbm > gpurun_out/r02y_bench_early.json 2> gpurun_out/r02y_bench_early.err
