mkdir -p gpurun_out/r02b
timeout 600 python -m pytest tests/test_gpu_glue.py -q -m gpu 2>&1 | tail -8
