# r02 pass 2: parity of the new epilogue / N-stationary paths, then A/B of the ResNet-50 step
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu2.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02_pytest_gpu2.log
timeout 300 python bench.py --no-cpu-baseline --layer-report gpurun_out/r02_layers1_resnet50.json > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r02_bench1.json
timeout 300 python bench.py --no-cpu-baseline --opt epi_split=0 --layer-report gpurun_out/r02_layers1_nosplit.json > gpurun_out/r02_bench1_nosplit.json 2> gpurun_out/r02_bench1_nosplit.err; echo "bench nosplit rc=$?"; cut -c1-200 gpurun_out/r02_bench1_nosplit.json
timeout 300 python bench.py --no-cpu-baseline --opt n_stationary=0 --layer-report gpurun_out/r02_layers1_nonstat.json > gpurun_out/r02_bench1_nonstat.json 2> gpurun_out/r02_bench1_nonstat.err; echo "bench nonstat rc=$?"; cut -c1-200 gpurun_out/r02_bench1_nonstat.json
timeout 300 python tools/trace_layer.py --layers l1.1.conv3,l2.1.conv3,l3.1.conv3,l4.1.conv3,l3.1.conv1,l1.1.conv2 --tiles 16 --skip 4 > gpurun_out/r02_trace1.txt 2>&1; echo "trace rc=$?"
