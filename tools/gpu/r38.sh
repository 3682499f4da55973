set -x
TAG=${TAG:-r38}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
run() { timeout 200 python bench.py --steps 100 --warmup 20 --no-secondary "$@" > gpurun_out/${TAG}_tmp.json 2>gpurun_out/${TAG}_b.err; python -c "
import json,sys;d=json.loads(open('gpurun_out/${TAG}_tmp.json').read().strip().splitlines()[-1]);print('bench','$*',d['ms_per_step'],d['clocks']['sm_mhz'],d['gpu_launches'])"; }
for rep in 1 2 3; do run; PSVAE_B200_LIB=$PWD/build/ab/libpsvae_prev.so run; done
timeout 300 python bench.py --steps 4 --warmup 3 --no-secondary > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 4 --warmup 3 --no-secondary > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu list rc=$?"
