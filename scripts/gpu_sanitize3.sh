mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
run() { tool=$1; tag=$2; shift 2
  timeout 1500 $SAN --tool $tool --error-exitcode 77 --log-file gpurun_out/san3_${tool}_${tag}.log python -m pytest -m gpu -q "$@" > gpurun_out/san3_${tool}_${tag}.out 2>&1
  echo "$tool $tag rc=$? | $(grep -E 'SUMMARY' gpurun_out/san3_${tool}_${tag}.log | tail -1) | $(tail -1 gpurun_out/san3_${tool}_${tag}.out)"
  grep -E "Race reported|Invalid|hazards\]" gpurun_out/san3_${tool}_${tag}.log | cut -c1-300 | sort | uniq -c | head -8
}
run racecheck hyb tests/test_hyb_gpu.py
run racecheck net tests/test_net_gpu.py
run memcheck hybnet tests/test_hyb_gpu.py tests/test_net_gpu.py
run synccheck hybnet tests/test_hyb_gpu.py tests/test_net_gpu.py
