set -x
TAG=${TAG:-r29}
run() { timeout 200 python bench.py --steps 100 --warmup 20 --no-secondary "$@" > gpurun_out/${TAG}_tmp.json 2>gpurun_out/${TAG}_b.err; python -c "
import json,sys;d=json.loads(open('gpurun_out/${TAG}_tmp.json').read().strip().splitlines()[-1]);print('bench','$*',d['ms_per_step'],d['clocks']['sm_mhz'])"; }
for rep in 1 2 3; do run; run --opt tc_bn_rounds=1; done
PSVAE_OPT_TC_BN_ROUNDS=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or twin or train_step" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest.log
