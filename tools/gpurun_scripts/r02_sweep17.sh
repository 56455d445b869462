set -u
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cta_pairs or many_tiles" > gpurun_out/r02_pytest_pairs17.log 2>&1; echo "pair tests rc=$? $(tail -1 gpurun_out/r02_pytest_pairs17.log)"
run() { echo "== $*"; timeout 120 python tools/run_layers.py --iters 5 "$@" 2>&1 | cut -c1-60,130-175,235-420; }
run --network resnet50 --layers l1.0.conv2
run --network resnet50 --layers l1.0.conv2 --opt cta_pairs=1
run --network vgg16 --layers conv1_2
run --network vgg16 --layers conv1_2 --opt cta_pairs=1
run --network resnet18 --layers l1.0.conv1
run --network resnet18 --layers l1.0.conv1 --opt cta_pairs=1
