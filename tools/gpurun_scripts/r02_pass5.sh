set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_networks.py -m gpu -x -q > gpurun_out/r02_pytest_gpu5.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_pytest_gpu5.log
timeout 300 python bench.py --no-cpu-baseline --layer-report gpurun_out/r02_layers4_resnet50.json > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r02_bench4.json
timeout 200 python tools/trace_layer.py --layers l1.1.conv3,conv1 --tiles 12 --skip 4 > gpurun_out/r02_trace2.txt 2>&1; echo "trace rc=$?"; head -5 gpurun_out/r02_trace2.txt | cut -c1-150
