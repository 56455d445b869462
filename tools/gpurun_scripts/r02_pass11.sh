set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "narrow or resident_filter_halves" > gpurun_out/r02_pytest_new11.log 2>&1; echo "new tests rc=$? $(tail -1 gpurun_out/r02_pytest_new11.log)"
timeout 300 python bench.py --no-cpu-baseline --layer-report gpurun_out/r02_layers11_resnet50.json > gpurun_out/r02_bench11.json 2> gpurun_out/r02_bench11.err; echo "bench rc=$? $(cut -c1-200 gpurun_out/r02_bench11.json)"
timeout 300 python bench.py --no-cpu-baseline --opt early_weights=0 > gpurun_out/r02_bench11_noearly.json 2> gpurun_out/r02_bench11_noearly.err; echo "bench noearly rc=$? $(cut -c1-200 gpurun_out/r02_bench11_noearly.json)"
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_bench11b.json 2> gpurun_out/r02_bench11b.err; echo "bench again rc=$? $(cut -c1-200 gpurun_out/r02_bench11b.json)"
timeout 300 python bench.py --no-cpu-baseline --opt early_weights=0 > gpurun_out/r02_bench11_noearly_b.json 2> gpurun_out/r02_bench11_noearly_b.err; echo "bench noearly again rc=$? $(cut -c1-200 gpurun_out/r02_bench11_noearly_b.json)"
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_gpu11.log 2>&1; echo "all tests rc=$? $(tail -1 gpurun_out/r02_pytest_gpu11.log)"
