#!/bin/bash
# 8 B200: BASELINE config 4 through the product driver (8 shots on 8 shot domains; then one shot over 8 GPUs as y-slabs),
# strong scaling of config 3, weak scaling of the north-star with the multi-GPU parity line
mkdir -p gpurun_out /tmp/cfg4/acq /tmp/cfg4/seismograms
cd /tmp/cfg4
cat > configuration.txt <<'CFG'
dimension=3D
equationType=viscoelastic
NX=768
NY=768
NZ=768
UseVariableGrid=0
useVariableFDoperators=0
useStencilMatrix=1
partitioning=1
NumShotDomains=8
DH=10
DT=0.8e-03
T=0.08
spatialFDorder=8
ModelRead=0
ModelFilename=model/model
fileFormat=2
numRelaxationMechanisms=2
relaxationFrequency=5
relaxationFrequency2=50
velocityP=3500
velocityS=2000
rho=2000
tauP=0.1
tauS=0.1
FreeSurface=1
DampingBoundary=2
BoundaryWidth=20
DampingCoeff=8.0
VMaxCPML=3500
CenterFrequencyCPML=10
NPower=4
SourceFilename=acq/sources
ReceiverFilename=acq/receiver
SeismogramFilename=seismograms/seismogram
initSourcesFromSU=0
initReceiverFromSU=0
SeismogramFormat=2
normalizeTraces=0
useReceiversPerShot=0
writeSource=0
seismoDT=0.8e-03
snapType=0
WavefieldFileName=wavefields/wavefield
tFirstSnapshot=0
tLastSnapshot=2
tIncSnapshot=0.1
verbose=0
CFG
{ echo "# sourceNo X Y Z type wType wShape fc amp tShift"; for s in 1 2 3 4 5 6 7 8; do echo "$s $((64 + 80 * s)) 1 384 3 1 1 10.0 1.0e6 0.0"; done; } > acq/sources.txt
{ echo "# X Y Z type"; for r in $(seq 0 63); do echo "$((128 + 8 * r)) 1 384 3"; done; } > acq/receiver.txt
S=$GRAFT_REPO_ROOT/wave-simulation_b200/host/Simulation
( time timeout 900 $S configuration.txt ) > $GRAFT_REPO_ROOT/gpurun_out/r02_cfg4_8shots_8gpus.log 2>&1
ls -la seismograms | head -12 >> $GRAFT_REPO_ROOT/gpurun_out/r02_cfg4_8shots_8gpus.log
sed -i 's/NumShotDomains=8/NumShotDomains=1/' configuration.txt
{ echo "# sourceNo X Y Z type wType wShape fc amp tShift"; echo "1 384 1 384 3 1 1 10.0 1.0e6 0.0"; } > acq/sources.txt
( time timeout 900 $S configuration.txt ) > $GRAFT_REPO_ROOT/gpurun_out/r02_cfg4_1shot_8gpus.log 2>&1
cd $GRAFT_REPO_ROOT
grep -E "shot domain|Finished|Total runtime|ERROR|real" gpurun_out/r02_cfg4_8shots_8gpus.log gpurun_out/r02_cfg4_1shot_8gpus.log | cut -c1-200
