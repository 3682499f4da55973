set -x
TAG=${TAG:-r11}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
for rep in 1 2 3; do
  timeout 200 python bench.py --steps 100 --warmup 20 --no-secondary > gpurun_out/${TAG}_b${rep}.json 2>gpurun_out/${TAG}_b.err; python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_b${rep}.json').read().strip().splitlines()[-1]);print('bench',d['ms_per_step'],d['clocks']['sm_mhz'])"
done
timeout 300 python tools/epi_probe.py > gpurun_out/${TAG}_probe.log 2>&1; cat gpurun_out/${TAG}_probe.log
timeout 200 python tools/sample_bench.py > gpurun_out/${TAG}_sample.log 2>&1; cat gpurun_out/${TAG}_sample.log
