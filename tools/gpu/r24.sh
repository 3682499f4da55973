set -x
TAG=${TAG:-r24}
run() { timeout 200 python bench.py --steps 100 --warmup 20 --no-secondary "$@" > gpurun_out/${TAG}_tmp.json 2>gpurun_out/${TAG}_b.err; python -c "
import json,sys;d=json.loads(open('gpurun_out/${TAG}_tmp.json').read().strip().splitlines()[-1]);print('bench','$*',d['ms_per_step'],d['clocks']['sm_mhz'])"; }
for rep in 1 2; do
run
run --opt tc_epi_groups_max_k=256
run --opt wgrad_split_cap=32
run --opt pdl=0
done
