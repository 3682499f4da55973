set -x
TAG=${TAG:-r22}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fast_mode or train_step or golden or twin" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
run() { timeout 200 python bench.py --steps 100 --warmup 20 --no-secondary "$@" > gpurun_out/${TAG}_tmp.json 2>gpurun_out/${TAG}_b.err; python -c "
import json,sys;d=json.loads(open('gpurun_out/${TAG}_tmp.json').read().strip().splitlines()[-1]);print('bench','$*',d['ms_per_step'],d['clocks']['sm_mhz'])"; }
for rep in 1 2 3; do run; done
timeout 300 python bench.py --steps 4 --warmup 3 --no-secondary > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:latent_bwd -c 6 --csv --log-file gpurun_out/${TAG}_lat.csv python bench.py --steps 4 --warmup 3 --no-secondary > gpurun_out/${TAG}_ncu.log 2>&1; grep latent gpurun_out/${TAG}_lat.csv | tail -3 | cut -c1-60,200-
