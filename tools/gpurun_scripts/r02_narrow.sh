set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "narrow" > gpurun_out/r02_pytest_narrow.log 2>&1; echo "narrow tests rc=$? $(tail -1 gpurun_out/r02_pytest_narrow.log)"
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_gpu7.log 2>&1; echo "all tests rc=$? $(tail -1 gpurun_out/r02_pytest_gpu7.log)"
timeout 300 python bench.py --no-cpu-baseline --layer-report gpurun_out/r02_layers7_resnet50.json > gpurun_out/r02_bench7.json 2> gpurun_out/r02_bench7.err; echo "bench rc=$? $(cut -c1-200 gpurun_out/r02_bench7.json)"
for n in resnet18 vgg16 mobilenet_v2; do
  timeout 300 python bench.py --no-cpu-baseline --network $n --layer-report gpurun_out/r02_layers7_${n}.json > gpurun_out/r02_bench7_${n}.json 2> gpurun_out/r02_bench7_${n}.err; echo "$n rc=$? $(cut -c1-170 gpurun_out/r02_bench7_${n}.json)"
done
for n in resnet50_full resnet18_full vgg16_full; do
  timeout 300 python bench.py --no-cpu-baseline --network $n --batch 128 > gpurun_out/r02_bench7_${n}.json 2> gpurun_out/r02_bench7_${n}.err; echo "$n rc=$? $(cut -c1-170 gpurun_out/r02_bench7_${n}.json)"; tail -2 gpurun_out/r02_bench7_${n}.err
done
timeout 300 python tools/trace_layer.py --network resnet50 --layers conv1,l1.0.conv1,l1.0.conv2 --tiles 16 > gpurun_out/r02_trace7.txt 2>&1; echo "trace rc=$?"
