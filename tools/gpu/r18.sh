set -x
TAG=${TAG:-r18}
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "data_parallel" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 100 --warmup 20 --no-secondary > gpurun_out/${TAG}_n2.json 2> gpurun_out/${TAG}_n2.err; echo "N=2 rc=$?"; tail -1 gpurun_out/${TAG}_n2.json | cut -c1-2500
grep -v "Warning\|^$\|\*\*\*\|OMP_NUM" gpurun_out/${TAG}_n2.err | tail -5
