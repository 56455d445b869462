set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_tail" > gpurun_out/r02_pytest_fused.log 2>&1; echo "fused tests rc=$?"; tail -25 gpurun_out/r02_pytest_fused.log
timeout 600 python -m pytest tests/test_gpu_networks.py tests/test_gpu_ops.py -m gpu -x -q > gpurun_out/r02_pytest_fused2.log 2>&1; echo "net tests rc=$?"; tail -5 gpurun_out/r02_pytest_fused2.log
timeout 300 python bench.py --no-cpu-baseline --layer-report gpurun_out/r02_layers6_resnet50.json > gpurun_out/r02_bench6.json 2> gpurun_out/r02_bench6.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r02_bench6.json; tail -3 gpurun_out/r02_bench6.err
timeout 300 python bench.py --no-cpu-baseline --opt fuse=0 > gpurun_out/r02_bench6_nofuse.json 2> gpurun_out/r02_bench6_nofuse.err; echo "bench nofuse rc=$?"; cut -c1-200 gpurun_out/r02_bench6_nofuse.json
