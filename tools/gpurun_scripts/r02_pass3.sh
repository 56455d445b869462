set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu3.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02_pytest_gpu3.log
timeout 300 python bench.py --no-cpu-baseline --layer-report gpurun_out/r02_layers2_resnet50.json > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r02_bench2.json
