set -x
python tools/step_sections.py cfg2 12 > gpurun_out/r02c_sections_cfg2.log 2>&1
python tools/step_sections.py cfg4 10 > gpurun_out/r02c_sections_cfg4.log 2>&1
python tools/gpu_l2gran.py > gpurun_out/r02c_l2gran.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02c_tests.log
