# 8-GPU box: the driver's scaling launches at N = 8, 4, 2, 1 (no secondary legs)
set -x
TAG=${TAG:-r32}
for n in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 100 --warmup 20 --no-secondary > gpurun_out/${TAG}_n${n}.json 2> gpurun_out/${TAG}_n${n}.err
  echo "N=$n rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_n${n}.json').read().strip().splitlines()[-1]);print('N',$n,d['ms_per_step'],d['value'],d.get('dp_parity'))"
done
timeout 200 python bench.py --steps 100 --warmup 20 --no-secondary > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err; python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_n1.json').read().strip().splitlines()[-1]);print('N',1,d['ms_per_step'],d['value'])"
