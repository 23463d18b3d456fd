# usage: bash scripts/gpu_net_ab.sh TAG "ENV1" ... -- network parity tests, then the bench's network sections (2048 replicas) once per setting
TAG=$1; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_net_gpu.py tests/test_hyb_gpu.py tests/test_itscp_env_gpu.py -x -q 2>&1 | tail -2
i=0
for e in "$@"; do
  env $e timeout 600 python scripts/net_prof.py --net-replicas 2048 > gpurun_out/${TAG}_net_$i.txt 2>&1
  echo "$e"; grep -o "'value': [0-9.e+]*\|'fwd_ms': [0-9.]*\|'bwd_ms': [0-9.]*\|'ms_per_batch': [0-9.]*" gpurun_out/${TAG}_net_$i.txt | tr '\n' ' '; echo
  i=$((i+1))
done
