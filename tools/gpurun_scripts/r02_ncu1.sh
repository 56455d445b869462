set -u
mkdir -p gpurun_out
L=conv1,l1.1.conv2,l1.1.conv3,l3.1.conv1
timeout 200 python tools/run_layers.py --network resnet50 --layers $L --iters 1 > gpurun_out/r02_ncu1_plain.log 2>&1 &&
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'igemm_i8_kernel' -o gpurun_out/r02_ncu1 -f python tools/run_layers.py --network resnet50 --layers $L --iters 1 > gpurun_out/r02_ncu1.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/r02_ncu1.log
