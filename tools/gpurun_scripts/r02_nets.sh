set -u
mkdir -p gpurun_out
for n in resnet18 vgg16 mobilenet_v2; do
  timeout 300 python bench.py --no-cpu-baseline --network $n --layer-report gpurun_out/r02_layers_${n}.json > gpurun_out/r02_bench_${n}.json 2> gpurun_out/r02_bench_${n}.err; echo "$n rc=$? $(cut -c1-170 gpurun_out/r02_bench_${n}.json)"
done
for n in resnet50_full resnet18_full vgg16_full; do
  timeout 300 python bench.py --no-cpu-baseline --network $n --batch 128 > gpurun_out/r02_bench_${n}.json 2> gpurun_out/r02_bench_${n}.err; echo "$n rc=$? $(cut -c1-170 gpurun_out/r02_bench_${n}.json)"; tail -2 gpurun_out/r02_bench_${n}.err
done
