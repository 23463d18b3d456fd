mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
run() { tool=$1; tag=$2; shift 2
  timeout 400 $SAN --tool $tool --error-exitcode 77 --log-file gpurun_out/san4_${tool}_${tag}.log python -m pytest -m gpu -q "$@" > gpurun_out/san4_${tool}_${tag}.out 2>&1
  echo "$tool $tag rc=$? | $(grep -E 'SUMMARY' gpurun_out/san4_${tool}_${tag}.log | tail -1) | $(tail -1 gpurun_out/san4_${tool}_${tag}.out)"
}
run synccheck arzidm tests/test_arz_gpu.py tests/test_idm_gpu.py
run racecheck idm tests/test_idm_gpu.py
run racecheck arz tests/test_arz_gpu.py
