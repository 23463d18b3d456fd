# usage: bash scripts/gpu_ab3.sh TAG "ENV1" "ENV2" ... -- the full-size bench (device-resident part only, parity check on) once per environment setting, no tests
TAG=$1; shift
mkdir -p gpurun_out
i=0
for e in "$@"; do
  env $e timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-net --no-drivers > gpurun_out/${TAG}_ab_$i.json 2>gpurun_out/${TAG}_ab_$i.err
  python -c "
import json; d=json.load(open('gpurun_out/${TAG}_ab_$i.json')); print('$e', '%.4g' % d['value'], {k: round(v,1) for k,v in d['phase_ms_per_step'].items()}, 'parity', d['parity_check'] and d['parity_check']['ok'])" || tail -3 gpurun_out/${TAG}_ab_$i.err
  i=$((i+1))
done
