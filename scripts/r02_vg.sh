#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_host_layer.py tests/test_abi.py -q -m gpu -k "variable_grid or abi or product_driver_on_gpu" 2>&1 | tail -15) > gpurun_out/r02_vg_tests.log 2>&1
cat gpurun_out/r02_vg_tests.log
