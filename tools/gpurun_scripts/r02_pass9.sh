set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "narrow or resident_filter_halves" > gpurun_out/r02_pytest_new9.log 2>&1; echo "new tests rc=$? $(tail -1 gpurun_out/r02_pytest_new9.log)"
timeout 300 python tools/trace_layer.py --network resnet50 --layers conv1 --tiles 24 --skip 200 > gpurun_out/r02_trace8_conv1.txt 2>&1; echo "trace rc=$?"
timeout 300 python tools/trace_layer.py --network resnet50 --layers l1.0.conv2,l2.1.conv2 --tiles 16 --skip 20 > gpurun_out/r02_trace8_l1.txt 2>&1; echo "trace rc=$?"
bash tools/gpurun_scripts/r02_sweep8.sh > gpurun_out/r02_sweep8.txt 2>&1; echo "sweep rc=$?"
run() { echo "== $*"; timeout 120 python tools/run_layers.py --iters 5 "$@" 2>&1 | cut -c1-60,130-175,235-420; }
( run --network resnet50 --layers l2.1.conv2; run --network resnet50 --layers l2.1.conv2 --opt resident_filter=2
  run --network vgg16 --layers conv2_2; run --network vgg16 --layers conv2_2 --opt resident_filter=2
  run --network resnet18 --layers l2.1.conv1; run --network resnet18 --layers l2.1.conv1 --opt resident_filter=2 ) > gpurun_out/r02_sweep9.txt 2>&1
timeout 300 python bench.py --no-cpu-baseline --layer-report gpurun_out/r02_layers9_resnet50.json > gpurun_out/r02_bench9.json 2> gpurun_out/r02_bench9.err; echo "bench rc=$? $(cut -c1-200 gpurun_out/r02_bench9.json)"
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_gpu9.log 2>&1; echo "all tests rc=$? $(tail -1 gpurun_out/r02_pytest_gpu9.log)"
