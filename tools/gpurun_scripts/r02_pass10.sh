set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "narrow or resident_filter_halves" > gpurun_out/r02_pytest_new10.log 2>&1; echo "new tests rc=$? $(tail -1 gpurun_out/r02_pytest_new10.log)"
run() { echo "== $*"; timeout 120 python tools/run_layers.py --iters 5 "$@" 2>&1 | cut -c1-60,130-175,235-420; }
L=l2.0.conv2,l2.1.conv2,l3.0.conv1,l3.1.conv1
( run --network resnet50 --layers $L; run --network resnet50 --layers $L --opt resident_filter=2 
  run --network resnet50 --layers $L --opt stage_bufs=1 ) > gpurun_out/r02_sweep10.txt 2>&1
timeout 300 python bench.py --no-cpu-baseline --layer-report gpurun_out/r02_layers10_resnet50.json > gpurun_out/r02_bench10.json 2> gpurun_out/r02_bench10.err; echo "bench rc=$? $(cut -c1-200 gpurun_out/r02_bench10.json)"
timeout 300 python bench.py --no-cpu-baseline --opt resident_filter=2 > gpurun_out/r02_bench10_nopairres.json 2> gpurun_out/r02_bench10_nopairres.err; echo "bench nopairres rc=$? $(cut -c1-200 gpurun_out/r02_bench10_nopairres.json)"
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_gpu10.log 2>&1; echo "all tests rc=$? $(tail -1 gpurun_out/r02_pytest_gpu10.log)"
