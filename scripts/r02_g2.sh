#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_host_layer.py tests/test_multigpu_nccl.py tests/test_gpu_long.py -q -m gpu 2>&1 | tail -25) > gpurun_out/r02_g2_tests.log 2>&1
cat gpurun_out/r02_g2_tests.log
