# racecheck / memcheck over more of the parity suite (small cases only; sanitizer slows kernels ~50x)
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
run() { tool=$1; tag=$2; shift 2
  timeout 1200 $SAN --tool $tool --error-exitcode 77 --log-file gpurun_out/san_${tool}_${tag}.log python -m pytest -m gpu -q "$@" > gpurun_out/san_${tool}_${tag}.out 2>&1
  echo "$tool $tag rc=$? | $(grep -E 'SUMMARY' gpurun_out/san_${tool}_${tag}.log | tail -1) | $(tail -1 gpurun_out/san_${tool}_${tag}.out)"
  grep -E "Race reported|Invalid|hazards\]" gpurun_out/san_${tool}_${tag}.log | cut -c1-400 | sort | uniq -c | head -12
}
run racecheck hyb tests/test_hyb_gpu.py
run racecheck net tests/test_net_gpu.py
run racecheck idm tests/test_idm_gpu.py tests/test_convert_gpu.py
run racecheck arz tests/test_arz_gpu.py
run memcheck arz tests/test_arz_gpu.py tests/test_idm_gpu.py tests/test_net_gpu.py
run memcheck env tests/test_itscp_env_gpu.py tests/test_inverse_gpu.py tests/test_dropin_gpu.py
