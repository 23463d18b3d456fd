# usage: bash scripts/gpu_ncu_arz.sh TAG K [env...]  -- ncu --set full of the two ARZ rollout kernels on a small batch
set -x
TAG=$1; K=$2
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"arz_rollout" -c 2 -o gpurun_out/${TAG} python bench.py --lanes 2368 --micro-lanes 1024 --sim-steps 64 --ckpt-every $K --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}.log 2>&1
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_sass.csv 2>/dev/null
ls -la gpurun_out | tail -5
