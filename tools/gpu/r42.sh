set -x
TAG=${TAG:-r42}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${TAG}_pytest.log
