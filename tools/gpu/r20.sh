set -x
TAG=${TAG:-r20}
timeout 600 python -m pytest tests/test_gpu_chain.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 200 python tools/sample_bench.py > gpurun_out/${TAG}_sample.log 2>&1; grep "chain=1" gpurun_out/${TAG}_sample.log
timeout 200 python tools/chain_trace.py > gpurun_out/${TAG}_chain_trace.log 2>&1; cat gpurun_out/${TAG}_chain_trace.log
