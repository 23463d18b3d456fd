# usage: bash scripts/gpu_ncu_net2.sh TAG [kernel regex] [launch count] -- ncu --set full of the network rollout kernels on 592 replicas, raw + per-source-line pages
TAG=$1; RE=${2:-net_rollout}; CNT=${3:-2}
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$RE" -c $CNT -o gpurun_out/${TAG} python scripts/net_prof.py > gpurun_out/${TAG}.log 2>&1
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}_cs.csv 2>/dev/null
python scripts/ncu_lines.py gpurun_out/${TAG}_cs.csv 40 > gpurun_out/${TAG}_lines.txt 2>&1
rm -f gpurun_out/${TAG}.ncu-rep
tail -3 gpurun_out/${TAG}.log
ls -la gpurun_out | grep ${TAG}
