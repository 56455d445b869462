set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02_pytest_gpu6.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02_pytest_gpu6.log
timeout 300 python bench.py --no-cpu-baseline --layer-report gpurun_out/r02_layers5_resnet50.json > gpurun_out/r02_bench5.json 2> gpurun_out/r02_bench5.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r02_bench5.json
L=conv1,l1.1.conv2,l1.1.conv1,l2.0.conv1
timeout 200 python tools/run_layers.py --network resnet50 --layers $L --iters 1 > gpurun_out/r02_ncu2_plain.log 2>&1 &&
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'igemm_i8_kernel' -o gpurun_out/r02_ncu2 -f python tools/run_layers.py --network resnet50 --layers $L --iters 1 > gpurun_out/r02_ncu2.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r02_ncu2.log
