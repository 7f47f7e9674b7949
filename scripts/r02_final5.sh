#!/bin/bash
# last GPU check of round 2: the whole -m gpu suite on the library as rebuilt in the final session (host-layer changes: mark-matrix
# receivers, common-offset profiles, trace gain, model file variants)
mkdir -p gpurun_out
(timeout 280 python -m pytest tests -q -x -m gpu 2>&1 | tail -6) > gpurun_out/r02_final5_tests.log 2>&1
cat gpurun_out/r02_final5_tests.log
