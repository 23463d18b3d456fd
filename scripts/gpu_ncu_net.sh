# usage: bash scripts/gpu_ncu_net.sh TAG -- ncu --set full of the network rollout kernels (macro + hybrid) on 592 replicas
TAG=$1
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"net_rollout|hyb_rollout" -c 4 -o gpurun_out/${TAG} python scripts/net_prof.py > gpurun_out/${TAG}.log 2>&1
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_sass.csv 2>/dev/null
tail -3 gpurun_out/${TAG}.log
ls -la gpurun_out | grep ${TAG}
