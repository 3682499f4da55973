# TMEM read microbenchmark + sampling chain trace + A/B of tc_epi_groups_max_k
set -x
TAG=${TAG:-r10}
timeout 120 build/tmem_ld_bench > gpurun_out/${TAG}_tmem.log 2>&1; echo "tmem rc=$?"; cat gpurun_out/${TAG}_tmem.log
timeout 200 python tools/chain_trace.py > gpurun_out/${TAG}_chain_trace.log 2>&1; echo "trace rc=$?"; cat gpurun_out/${TAG}_chain_trace.log
for rep in 1 2; do
for k in 128 256 512; do
  timeout 200 python bench.py --steps 100 --warmup 20 --no-secondary --opt tc_epi_groups_max_k=$k > gpurun_out/${TAG}_k${k}_${rep}.json 2>gpurun_out/${TAG}_k${k}.err; python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_k${k}_${rep}.json').read().strip().splitlines()[-1]);print('maxk',$k,d['ms_per_step'],d['clocks']['sm_mhz'])"
done; done
