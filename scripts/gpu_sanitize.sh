# usage: bash scripts/gpu_sanitize.sh  -- compute-sanitizer memcheck / racecheck / synccheck over the smoke run and small parity tests
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SAN=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  timeout 600 $SAN --tool $tool --error-exitcode 77 --log-file gpurun_out/san_${tool}_smoke.log python __graft_entry__.py --smoke > gpurun_out/san_${tool}_smoke.out 2>&1
  echo "$tool smoke rc=$? $(grep -c 'ERROR SUMMARY' gpurun_out/san_${tool}_smoke.log) $(grep 'ERROR SUMMARY' gpurun_out/san_${tool}_smoke.log | tail -1)"
done
for tool in memcheck racecheck; do
  timeout 900 $SAN --tool $tool --error-exitcode 77 --log-file gpurun_out/san_${tool}_tests.log python -m pytest -m gpu -x -q \
     tests/test_hyb_gpu.py::test_plain_mode_chain_matches_live_reference tests/test_net_gpu.py::test_itscp_macro_matches_live_reference \
     "tests/test_hyb_gpu.py::test_hybrid_itscp_matches_live_reference[h]" tests/test_convert_gpu.py > gpurun_out/san_${tool}_tests.out 2>&1
  echo "$tool tests rc=$? $(grep 'ERROR SUMMARY' gpurun_out/san_${tool}_tests.log | tail -1) $(tail -1 gpurun_out/san_${tool}_tests.out)"
done
