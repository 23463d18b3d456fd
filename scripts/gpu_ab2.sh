# usage: bash scripts/gpu_ab2.sh TAG "ENV1" "ENV2" ... -- the whole GPU suite, then the full-size bench (device-resident part only) once per environment setting
set -x
TAG=$1; shift
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
i=0
for e in "$@"; do
  env $e python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-net --no-drivers > gpurun_out/${TAG}_ab_$i.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/${TAG}_ab_$i.json')); print('$e', '%.4g' % d['value'], {k: round(v,1) for k,v in d['phase_ms_per_step'].items()}, 'parity', d['parity_check'] and d['parity_check']['ok'])"
  i=$((i+1))
done
