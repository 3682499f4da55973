set -x
TAG=${TAG:-r30}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fast_mode" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest.log
run() { timeout 200 python bench.py --steps 100 --warmup 20 --no-secondary "$@" > gpurun_out/${TAG}_tmp.json 2>gpurun_out/${TAG}_b.err; python -c "
import json,sys;d=json.loads(open('gpurun_out/${TAG}_tmp.json').read().strip().splitlines()[-1]);print('bench','$*',d['ms_per_step'],d['clocks']['sm_mhz'])"; }
for rep in 1 2; do
run
run --opt wgrad_lpt=1
run --opt wgrad_lpt=1 --opt wgrad_splits=9
run --opt wgrad_lpt=1 --opt wgrad_splits=12
run --opt wgrad_lpt=1 --opt wgrad_splits=8
run --opt wgrad_lpt=1 --opt wgrad_splits=15
done
