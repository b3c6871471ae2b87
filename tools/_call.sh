set -x
B="python bench.py --warmup 5 --no-cpu-baseline --no-hbm"
for i in 1 2; do DCL_DEBUG_PY_TIMES=1 DCL_BENCH_DUMP_STEPS=1 $B --steps 100 --workload cfg4 > gpurun_out/r02u_$i.json 2> gpurun_out/r02u_$i.err; done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err
grep -E "slow step|look-ahead" gpurun_out/r02u_*.err
