# final_pass.sh — the measurement calls behind profiles/ (run each numbered block as ONE gpurun call; at most one ncu per call,
# and only after the same command has exited 0 without it).
#   bash tools/gpurun_scripts/final_pass.sh 1      tests, bench line, per-layer tables of the four networks, ncu launch list of bench.py
#   bash tools/gpurun_scripts/final_pass.sh 2|3|4  ncu --set full over representative layers of resnet50 | vgg16 | mobilenet_v2
set -u
case "$1" in
1)
  timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/pytest_gpu.log
  timeout 400 python bench.py --layer-report gpurun_out/final_layers_resnet50.json > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
  cut -c1-200 gpurun_out/final_bench.json
  for n in vgg16 resnet18 mobilenet_v2; do
    timeout 200 python bench.py --no-cpu-baseline --network $n --layer-report gpurun_out/final_layers_${n}.json > gpurun_out/final_bench_${n}.json 2>/dev/null
    cut -c1-160 gpurun_out/final_bench_${n}.json
  done
  timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b21.json 2> gpurun_out/b21.err &&
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
      --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  echo "ncu rc=$?"; wc -l gpurun_out/r01_bench_launches.csv ;;
2) NET=resnet50; L=conv1,l1.1.conv1,l1.1.conv2,l1.1.conv3,l2.0.conv1,l2.1.conv2,l2.1.conv3,l3.1.conv1,l3.1.conv2,l3.1.conv3,l4.1.conv2; K='igemm_i8_kernel|stem_xform' ;;
3) NET=vgg16; L=; K='igemm_i8_kernel|stem_xform' ;;
4) NET=mobilenet_v2; L=stem,b0.dw,b0.project,b1.expand,b1.dw,b2.dw,b3.dw,b4.expand,b4.dw,b7.dw,b7.project,b14.expand,b14.dw,last; K='igemm_i8_kernel|stem_xform|depthwise' ;;
esac
if [ "$1" != 1 ]; then
  timeout 200 python tools/run_layers.py --network $NET --layers "$L" --iters 1 > gpurun_out/final_rl_$NET.log 2>&1 &&
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:"$K" -o gpurun_out/r01_final_$NET -f \
      python tools/run_layers.py --network $NET --layers "$L" --iters 1 > gpurun_out/ncu_$NET.log 2>&1
  echo "rc=$?"; tail -1 gpurun_out/ncu_$NET.log
fi
