set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r03a_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r03a_bench.json 2> gpurun_out/r03a_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r03a_bench_ref.json 2> gpurun_out/r03a_bench_ref.err
python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/r03a_smoke.log 2>&1
