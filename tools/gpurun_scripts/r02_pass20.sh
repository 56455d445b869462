set -u
mkdir -p gpurun_out
for i in 1 2; do timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_bench20_$i.json 2> gpurun_out/r02_bench20_$i.err; echo "bench $i rc=$? $(cut -c1-180 gpurun_out/r02_bench20_$i.json)"; done
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_gpu20.log 2>&1; echo "all tests rc=$? $(tail -1 gpurun_out/r02_pytest_gpu20.log)"
timeout 300 python bench.py --no-cpu-baseline --batch 64 | cut -c1-180
