set -u
mkdir -p gpurun_out
timeout 300 python tools/trace_layer.py --network resnet50 --layers conv1 --tiles 24 --skip 200 > gpurun_out/r02_trace8_conv1.txt 2>&1; echo "trace rc=$?"
timeout 300 python tools/trace_layer.py --network resnet50 --layers l1.0.conv2,l1.0.conv1 --tiles 16 --skip 60 > gpurun_out/r02_trace8_l1.txt 2>&1; echo "trace rc=$?"
timeout 300 python tools/trace_layer.py --network vgg16 --layers conv1_2 --tiles 16 --skip 200 > gpurun_out/r02_trace8_vgg.txt 2>&1; echo "trace rc=$?"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "narrow" > gpurun_out/r02_pytest_narrow.log 2>&1; echo "narrow tests rc=$? $(tail -1 gpurun_out/r02_pytest_narrow.log)"
bash tools/gpurun_scripts/r02_sweep8.sh > gpurun_out/r02_sweep8.txt 2>&1; echo "sweep rc=$?"
