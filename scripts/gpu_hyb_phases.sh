# usage: bash scripts/gpu_hyb_phases.sh TAG -- rebuilds the library with -DDHTS_PHASE_TIMING ON THE BOX (the in-tree .so of the snapshot is replaced there only) and prints the per-phase cycles of a config-4 episode
TAG=$1
mkdir -p gpurun_out
DHTS_NVCC_EXTRA=-DDHTS_PHASE_TIMING python -c "
import sys; sys.path.insert(0, 'diff-hybrid-traffic-sim_b200')
import _build; _build.build(force=True)" > gpurun_out/${TAG}_build.log 2>&1
python scripts/hyb_phases.py > gpurun_out/${TAG}_phases.txt 2>&1
cat gpurun_out/${TAG}_phases.txt | tail -30
