# usage: bash tools/scale_run.sh N [exchange]   (under gpurun --gpus N)
N=$1; EX=${2:-peer}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --exchange $EX --no-parity --no-latency \
  > gpurun_out/r2_bench_${N}gpu_${EX}.json 2> gpurun_out/r2_bench_${N}gpu_${EX}.err
tail -c 1500 gpurun_out/r2_bench_${N}gpu_${EX}.json
tail -3 gpurun_out/r2_bench_${N}gpu_${EX}.err
