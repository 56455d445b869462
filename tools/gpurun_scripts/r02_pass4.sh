set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_networks.py -m gpu -x -q > gpurun_out/r02_pytest_gpu4.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_pytest_gpu4.log
L=l1.1.conv3,l2.1.conv3,l3.1.conv3,l2.0.downsample,l1.0.downsample
run() { echo "== $*"; timeout 120 python tools/run_layers.py --network resnet50 --layers $L --iters 5 "$@" 2>&1 | cut -c1-60,150-420; }
run
run --opt stage_bufs=1
timeout 300 python bench.py --no-cpu-baseline --layer-report gpurun_out/r02_layers3_resnet50.json > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r02_bench3.json
timeout 300 python bench.py --no-cpu-baseline --opt stage_bufs=1 > gpurun_out/r02_bench3_b1.json 2> gpurun_out/r02_bench3_b1.err; echo "bench bufs1 rc=$?"; cut -c1-200 gpurun_out/r02_bench3_b1.json
