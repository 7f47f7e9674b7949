#!/bin/bash
# One GPU-box session: smoke, GPU parity tests, bench, ncu launch list, ncu full capture. Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
cp MEASURED_PEAKS.json gpurun_out/ 2>/dev/null
if [ -z "$SKIP_SMOKE" ]; then
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
fi
if [ -z "$SKIP_TESTS" ]; then
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
fi
if [ -z "$SKIP_BENCH" ]; then
echo "== bench"; timeout 900 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -3 gpurun_out/bench.log
fi
if [ -z "$SKIP_NCU" ]; then
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --nx 512 --ny 512 --nz 512 ${BENCH_ARGS} > gpurun_out/bench_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/bench_ncu.log
echo "== ncu full"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-kFast} -s 8 -c 4 -f -o gpurun_out/prof python bench.py --steps 2 --warmup 3 --no-cpu --nx 512 --ny 512 --nz 512 ${BENCH_ARGS} > gpurun_out/bench_ncufull.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/bench_ncufull.log
fi
