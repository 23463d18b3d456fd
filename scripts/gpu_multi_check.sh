set -x
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/multi_${N}.json 2> gpurun_out/multi_${N}.err
tail -3 gpurun_out/multi_${N}.err
cat gpurun_out/multi_${N}.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/multi_ref_${N}.json 2> gpurun_out/multi_ref_${N}.err
cat gpurun_out/multi_ref_${N}.json
