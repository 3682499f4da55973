set -x
TAG=${TAG:-r41}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 100 --warmup 20 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 4 --warmup 3 --no-secondary > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 4 --warmup 3 --no-secondary > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu list rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
