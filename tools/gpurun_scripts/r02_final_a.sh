set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r02f_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $O/r02f_smoke.log)"
timeout 600 python bench.py --layer-report $O/r02f_layers_resnet50.json > $O/r02f_bench.json 2> $O/r02f_bench.err; echo "bench rc=$? $(cut -c1-160 $O/r02f_bench.json)"
timeout 600 python bench.py --impl reference > $O/r02f_bench_reference.json 2> $O/r02f_bench_reference.err; echo "reference rc=$? $(cut -c1-200 $O/r02f_bench_reference.json)"
for n in resnet18 vgg16 mobilenet_v2; do
  timeout 300 python bench.py --no-cpu-baseline --network $n --layer-report $O/r02f_layers_${n}.json > $O/r02f_bench_${n}.json 2> $O/r02f_bench_${n}.err; echo "$n rc=$? $(cut -c1-170 $O/r02f_bench_${n}.json)"
done
for n in resnet50_full resnet18_full vgg16_full; do
  timeout 300 python bench.py --no-cpu-baseline --network $n --batch 128 > $O/r02f_bench_${n}.json 2> $O/r02f_bench_${n}.err; echo "$n rc=$? $(cut -c1-170 $O/r02f_bench_${n}.json)"
done
timeout 300 python bench.py --no-cpu-baseline --scaling strong > $O/r02f_bench_strong1.json 2> $O/r02f_bench_strong1.err; echo "strong1 rc=$?"
timeout 300 python bench.py --impl reference-gpu > $O/r02f_reference_gpu.json 2> $O/r02f_reference_gpu.err; echo "reference-gpu rc=$? $(cut -c1-120 $O/r02f_reference_gpu.json)"
timeout 300 python tools/trace_layer.py --network resnet50 --layers conv1,l1.1.conv2,l1.1.conv3,l2.1.conv2,l3.1.conv1,l3.1.conv2,l3.1.conv3,l4.1.conv3 --tiles 12 --skip 4 > $O/r02f_trace.txt 2>&1; echo "trace rc=$?"
( cd lowbitdnn-project_b200/cpp && timeout 300 build/check 5 10 > ../../$O/r02f_cpp_check.log 2>&1; echo "check rc=$?"; timeout 300 build/int8_bench --network resnet50 --repeats 5 > ../../$O/r02f_cpp_int8_bench_resnet50.log 2>&1; echo "int8_bench rc=$?" )
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/r02f_bench_launches.csv python bench.py --no-cpu-baseline --steps 2 --warmup 1 > $O/r02f_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
