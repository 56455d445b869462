set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tail_split" > gpurun_out/r02_pytest_tail21.log 2>&1; echo "tail tests rc=$? $(tail -1 gpurun_out/r02_pytest_tail21.log)"
run() { echo "== $*"; timeout 120 python tools/run_layers.py --network resnet50 --iters 5 "$@" 2>&1 | cut -c1-60,130-175,235-420; }
L=l3.0.conv2,l3.1.conv1,l3.1.conv2
( run --layers $L; run --layers $L --opt tail_split=0 ) > gpurun_out/r02_sweep21.txt 2>&1
for i in 1 2; do timeout 300 python bench.py --no-cpu-baseline --layer-report gpurun_out/r02_layers21_resnet50.json > gpurun_out/r02_bench21_$i.json 2> gpurun_out/r02_bench21_$i.err; echo "bench $i rc=$? $(cut -c1-180 gpurun_out/r02_bench21_$i.json)"; done
timeout 300 python bench.py --no-cpu-baseline --opt tail_split=0 > gpurun_out/r02_bench21_off.json 2> gpurun_out/r02_bench21_off.err; echo "bench off rc=$? $(cut -c1-180 gpurun_out/r02_bench21_off.json)"
