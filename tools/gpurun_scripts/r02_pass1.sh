# r02 pass 1: full GPU test suite, the C++ boundary on hardware, reference yardsticks, pipeline traces, bench line
set -u
mkdir -p gpurun_out
( timeout 900 python bench.py --impl reference --ref-full --steps 3 --warmup 1 > gpurun_out/r02_reference_full.json 2> gpurun_out/r02_reference_full.err; echo "ref-full rc=$?" ) &
REFPID=$!
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02_pytest_gpu.log
cd lowbitdnn-project_b200/cpp
( timeout 300 build/check 5 10 ) > ../../gpurun_out/r02_cpp_check.log 2>&1; echo "check rc=$?"; tail -4 ../../gpurun_out/r02_cpp_check.log
( timeout 300 build/int8_bench --network resnet50 --repeats 5 ) > ../../gpurun_out/r02_cpp_int8_bench.log 2>&1; echo "int8_bench rc=$?"; tail -3 ../../gpurun_out/r02_cpp_int8_bench.log
( timeout 300 build/benchmark_app apps/config.json ../../gpurun_out/r02_cpp_output.json --limit 24 ) > ../../gpurun_out/r02_cpp_benchmark_app.log 2>&1; echo "benchmark_app rc=$?"; tail -2 ../../gpurun_out/r02_cpp_benchmark_app.log
cd ../..
wait $REFPID
timeout 300 python bench.py --impl reference-gpu > gpurun_out/r02_reference_gpu.json 2> gpurun_out/r02_reference_gpu.err; echo "reference-gpu rc=$?"; cut -c1-600 gpurun_out/r02_reference_gpu.json
timeout 300 python tools/trace_layer.py --layers conv1,l1.1.conv2,l1.1.conv1,l2.1.conv3,l2.1.conv2,l3.0.conv1,l3.1.conv1,l3.1.conv3,l4.1.conv1,l4.1.conv3,l4.1.conv2 --tiles 16 --skip 4 > gpurun_out/r02_trace0.txt 2>&1; echo "trace rc=$?"
timeout 400 python bench.py --layer-report gpurun_out/r02_layers0_resnet50.json > gpurun_out/r02_bench0.json 2> gpurun_out/r02_bench0.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/r02_bench0.json
timeout 300 python bench.py --scaling strong --no-cpu-baseline > gpurun_out/r02_bench0_strong1.json 2> gpurun_out/r02_bench0_strong1.err; echo "bench strong rc=$?"
