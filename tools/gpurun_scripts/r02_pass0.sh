# r02 pass 0: the C++ boundary on hardware + pipeline traces of the layers furthest below roofline
set -u
cd lowbitdnn-project_b200
( timeout 300 cpp/build/check 5 10 ) > ../gpurun_out/r02_cpp_check.log 2>&1; echo "check rc=$?"; tail -4 ../gpurun_out/r02_cpp_check.log
( timeout 300 cpp/build/int8_bench --network resnet50 --repeats 5 ) > ../gpurun_out/r02_cpp_int8_bench.log 2>&1; echo "int8_bench rc=$?"; tail -3 ../gpurun_out/r02_cpp_int8_bench.log
( timeout 300 cpp/build/benchmark_app cpp/apps/config.json ../gpurun_out/r02_cpp_output.json --limit 24 ) > ../gpurun_out/r02_cpp_benchmark_app.log 2>&1; echo "benchmark_app rc=$?"; tail -3 ../gpurun_out/r02_cpp_benchmark_app.log
cd ..
timeout 300 python tools/trace_layer.py --layers conv1,l1.1.conv2,l1.1.conv1,l2.1.conv3,l2.1.conv2,l3.0.conv1,l3.1.conv1,l3.1.conv3,l4.1.conv1,l4.1.conv3,l4.1.conv2 --tiles 16 --skip 4 > gpurun_out/r02_trace0.txt 2>&1; echo "trace rc=$?"
timeout 400 python bench.py --layer-report gpurun_out/r02_layers0_resnet50.json > gpurun_out/r02_bench0.json 2> gpurun_out/r02_bench0.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/r02_bench0.json
