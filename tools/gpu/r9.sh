# round 2, session 2, call 1: tests + bench of the cleaned build (decode_chain on by default) + launch list with DRAM bytes + one ncu --set full pass
set -x
TAG=${TAG:-r9}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 100 --warmup 20 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --steps 4 --warmup 3 --no-secondary > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 4 --warmup 3 --no-secondary > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -s 80 -c 18 -o gpurun_out/${TAG}_full python bench.py --steps 4 --warmup 3 --no-secondary > gpurun_out/${TAG}_ncufull.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/${TAG}_*
