set -u
mkdir -p gpurun_out
for i in 1 2; do timeout 300 python bench.py --no-cpu-baseline --layer-report gpurun_out/r02_layers15_resnet50.json > gpurun_out/r02_bench15_$i.json 2> gpurun_out/r02_bench15_$i.err; echo "bench $i rc=$? $(cut -c1-180 gpurun_out/r02_bench15_$i.json)"; done
timeout 900 python -m pytest tests/test_gpu_networks.py tests/test_gpu_ops.py -x -q -m gpu > gpurun_out/r02_pytest_net15.log 2>&1; echo "net tests rc=$? $(tail -1 gpurun_out/r02_pytest_net15.log)"
