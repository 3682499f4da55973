# usage (8-GPU box): bash tools/gpu/run_scaling.sh   -- the driver's scaling launch at N = 2, 4, 8 (no secondary legs) + N = 8 sampling
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 200 --warmup 20 --no-secondary > gpurun_out/bench_n${n}_${TAG:-x}.log 2> gpurun_out/bench_n${n}_${TAG:-x}.err
  echo "N=$n rc=$?"; tail -1 gpurun_out/bench_n${n}_${TAG:-x}.log | cut -c1-260
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/bench_n8_full_${TAG:-x}.log 2> gpurun_out/bench_n8_full_${TAG:-x}.err
echo "N=8 full rc=$?"; tail -1 gpurun_out/bench_n8_full_${TAG:-x}.log | cut -c1-200
