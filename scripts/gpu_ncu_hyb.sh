TAG=$1
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"hyb_rollout" -c 2 -o gpurun_out/${TAG} python scripts/c4_breakdown.py > gpurun_out/${TAG}.log 2>&1
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}_cs.csv 2>/dev/null
ls -la gpurun_out | grep ${TAG}
