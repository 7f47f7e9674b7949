cp wave-simulation_b200/csrc/libwavesim_cuda.so /tmp/new.so
echo "== new"; WS_MARCH_LANES=1 WS_MARCH_SINGLE=0 WORKLOADS="northstar cfg4" STAGES="0" scripts/march_sweep.sh
cp alt_old.so wave-simulation_b200/csrc/libwavesim_cuda.so
echo "== old"; WS_MARCH_LANES=1 WORKLOADS="northstar cfg4" STAGES="0" scripts/march_sweep.sh
cp /tmp/new.so wave-simulation_b200/csrc/libwavesim_cuda.so
echo "== new again"; WS_MARCH_LANES=1 WS_MARCH_SINGLE=0 WORKLOADS="northstar" STAGES="0" scripts/march_sweep.sh
