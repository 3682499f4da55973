# usage: TAG=name bash tools/gpu/run_all.sh   -- probe + trace + GPU tests + bench + ncu launch list
T=${TAG:-x}
timeout 300 python tools/epi_probe.py > gpurun_out/epi_probe_$T.log 2>&1; echo "probe rc=$?"; tail -40 gpurun_out/epi_probe_$T.log
timeout 300 python tools/tc_trace.py > gpurun_out/tc_trace_$T.log 2>&1; echo "trace rc=$?"; head -4 gpurun_out/tc_trace_$T.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_$T.log
timeout 300 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_$T.log 2>&1; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$T.log').read().strip().splitlines()[-1])
    print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], {k:v['value'] for k,v in d['secondary'].items()})
except Exception as e:
    print('bench parse fail', e); print(open('gpurun_out/bench_$T.log').read()[-2000:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 2 --warmup 3 --no-secondary > gpurun_out/ncu_$T.log 2>&1; echo "ncu rc=$?"
