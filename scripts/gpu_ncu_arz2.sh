# usage: bash scripts/gpu_ncu_arz2.sh TAG -- ncu --set full (with SASS source) of one bench-shaped lane chunk of the ARZ rollout kernels
TAG=${1:-r2z}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"arz_rollout" -c 2 -o gpurun_out/${TAG}_full python bench.py --lanes 6560 --micro-lanes 65536 --sim-steps 256 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-net --no-drivers --no-parity > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_full.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_full_sass.csv 2>/dev/null
rm -f gpurun_out/${TAG}_full.ncu-rep
ls -la gpurun_out | tail -4
