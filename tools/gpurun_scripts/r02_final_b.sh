set -u
mkdir -p gpurun_out
L=conv1,l1.1.conv2,l1.1.conv3,l2.1.conv2,l3.1.conv1,l3.1.conv2,l3.1.conv3,l4.1.conv3
timeout 200 python tools/run_layers.py --network resnet50 --layers $L --iters 1 > gpurun_out/r02f_ncu_plain.log 2>&1 &&
timeout 1500 ncu --set full --import-source on --clock-control none -k regex:'igemm_i8_kernel' -o gpurun_out/r02f_ncu -f python tools/run_layers.py --network resnet50 --layers $L --iters 1 > gpurun_out/r02f_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r02f_ncu.log
