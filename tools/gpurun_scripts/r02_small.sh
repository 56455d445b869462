set -u
timeout 300 python bench.py --no-cpu-baseline --batch 64 --layer-report gpurun_out/r02_layers_b64.json > gpurun_out/r02_bench_b64.json 2> gpurun_out/r02_bench_b64.err; echo "b64 rc=$?"; cut -c1-200 gpurun_out/r02_bench_b64.json
timeout 300 python bench.py --no-cpu-baseline --batch 64 --opt max_bn=128 --layer-report gpurun_out/r02_layers_b64_bn128.json > gpurun_out/r02_bench_b64_bn128.json 2> gpurun_out/r02_bench_b64_bn128.err; echo "b64 bn128 rc=$?"; cut -c1-200 gpurun_out/r02_bench_b64_bn128.json
timeout 300 python bench.py --no-cpu-baseline --batch 64 --opt cta_pairs=0 --layer-report gpurun_out/r02_layers_b64_nopair.json > gpurun_out/r02_bench_b64_nopair.json 2> gpurun_out/r02_bench_b64_nopair.err; echo "b64 nopair rc=$?"; cut -c1-200 gpurun_out/r02_bench_b64_nopair.json
