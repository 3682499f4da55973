set -x
TAG=${TAG:-r12}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fast_mode_engine_variants" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
run() { timeout 200 python bench.py --steps 100 --warmup 20 --no-secondary "$@" > gpurun_out/${TAG}_tmp.json 2>gpurun_out/${TAG}_b.err; python -c "
import json,sys;d=json.loads(open('gpurun_out/${TAG}_tmp.json').read().strip().splitlines()[-1]);print('bench','$*',d['ms_per_step'],d['clocks']['sm_mhz'])"; }
for rep in 1 2; do
run
run --opt wgrad_order=1
run --opt wgrad_order=2
run --opt tc_thin_two_per_sm=1
run --opt tc_thin_two_per_sm=1 --opt wgrad_order=2
done
