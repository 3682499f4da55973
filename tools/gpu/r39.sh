# ncu --set full of the two sampling kernels (the chained decoder and the one-launch Langevin loop), each after its own plain run
set -x
TAG=${TAG:-r39}
timeout 200 python tools/chain_trace.py 757760 > gpurun_out/${TAG}_chain_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decoder_chain -s 3 -c 1 -o gpurun_out/${TAG}_chain python tools/chain_trace.py 757760 > gpurun_out/${TAG}_chain_ncu.log 2>&1; echo "chain rc=$?"
timeout 200 python tools/langevin_bench.py 262144 > gpurun_out/${TAG}_lang_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:langevin -s 2 -c 1 -o gpurun_out/${TAG}_lang python tools/langevin_bench.py 262144 > gpurun_out/${TAG}_lang_ncu.log 2>&1; echo "lang rc=$?"
ls -la gpurun_out/${TAG}_*
