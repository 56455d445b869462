set -u
mkdir -p gpurun_out
ALL=conv1,l1.0.conv1,l1.0.conv2,l1.0.conv3,l1.0.downsample,l1.1.conv1,l1.1.conv2,l1.1.conv3,l1.2.conv1,l1.2.conv2,l1.2.conv3,l2.0.conv1,l2.0.conv2,l2.0.conv3,l2.0.downsample,l2.1.conv1,l2.1.conv2,l2.1.conv3,l2.2.conv1,l2.2.conv2,l2.2.conv3,l2.3.conv1,l2.3.conv2,l2.3.conv3,l3.0.conv1,l3.0.conv2,l3.0.conv3,l3.0.downsample,l3.1.conv1,l3.1.conv2,l3.1.conv3,l3.2.conv1,l3.2.conv2,l3.2.conv3,l3.3.conv1,l3.3.conv2,l3.3.conv3,l3.4.conv1,l3.4.conv2,l3.4.conv3,l3.5.conv1,l3.5.conv2,l3.5.conv3,l4.0.conv1,l4.0.conv2,l4.0.conv3,l4.0.downsample,l4.1.conv1,l4.1.conv2,l4.1.conv3,l4.2.conv1,l4.2.conv2,l4.2.conv3
timeout 600 python tools/trace_layer.py --network resnet50 --layers $ALL --tiles 2 --skip 0 2>&1 | grep "per-CTA\|^== " | cut -c1-150 > gpurun_out/r02_cta_times.txt; echo "trace rc=$?"
timeout 300 python bench.py --no-cpu-baseline --opt pdl=0 > gpurun_out/r02_bench_nopdl.json 2> gpurun_out/r02_bench_nopdl.err; echo "nopdl rc=$? $(cut -c1-200 gpurun_out/r02_bench_nopdl.json)"
timeout 300 python bench.py --no-cpu-baseline --batch 1024 > gpurun_out/r02_bench_b1024.json 2> gpurun_out/r02_bench_b1024.err; echo "b1024 rc=$? $(cut -c1-200 gpurun_out/r02_bench_b1024.json)"
timeout 300 python bench.py --no-cpu-baseline --batch 256 > gpurun_out/r02_bench_b256.json 2> gpurun_out/r02_bench_b256.err; echo "b256 rc=$? $(cut -c1-200 gpurun_out/r02_bench_b256.json)"
