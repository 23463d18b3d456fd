# usage: bash scripts/gpu_r2b.sh TAG -- whole GPU suite + episode times of the unchanged drivers (deferred / immediate)
set -x
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
for p in macro micro hybrid; do
  timeout 300 python baseline/run_drivers.py --impl dropin --problem $p --episodes 10 2>/dev/null | cut -c1-330
  timeout 300 python baseline/run_drivers.py --impl dropin --problem $p --episodes 4 --no-defer 2>/dev/null | cut -c1-200
done > gpurun_out/${TAG}_drivers.txt 2>&1
cat gpurun_out/${TAG}_drivers.txt
