set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_networks.py tests/test_gpu_ops.py -x -q -m gpu > gpurun_out/r02_pytest_net22.log 2>&1; echo "net tests rc=$? $(tail -1 gpurun_out/r02_pytest_net22.log)"
for i in 1 2; do timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_bench22_$i.json 2> gpurun_out/r02_bench22_$i.err; echo "bench $i rc=$? $(cut -c1-180 gpurun_out/r02_bench22_$i.json)"; done
timeout 300 python bench.py --no-cpu-baseline --opt early_weights=0 > gpurun_out/r02_bench22_off.json 2> gpurun_out/r02_bench22_off.err; echo "bench early off rc=$? $(cut -c1-180 gpurun_out/r02_bench22_off.json)"
for n in vgg16 resnet18; do timeout 300 python bench.py --no-cpu-baseline --network $n 2>/dev/null | cut -c1-170; done
