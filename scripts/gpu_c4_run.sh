# run_itscp_hybrid.sh (BASELINE.json configs[3]) through the headless env: 3 problems x 100 epochs, as the reference script
mkdir -p gpurun_out
for P in 1 2 3; do
  T0=$(date +%s.%N)
  timeout 900 python -m dhts_b200.run_itscp --mode=hybrid --problem=$P --n_trial=1 --n_intersection=3 --n_lane=1 --lane_length=5 --speed_limit=60 --simulation_length=20 --signal_length=4 --n_episode=100 --lr=1e-4 --seed $((10+P)) --out gpurun_out/itscp_hybrid_p$P > gpurun_out/itscp_hybrid_p$P.log 2>&1
  echo "problem $P wall $(python -c "import time; print(round(time.time()-$T0,1))") s"
  head -2 gpurun_out/itscp_hybrid_p$P.log | cut -c1-80; grep "epoch 100" gpurun_out/itscp_hybrid_p$P.log
  cat gpurun_out/itscp_hybrid_p$P/trial_0/eval.txt | tr '\n' ' '; echo
done
