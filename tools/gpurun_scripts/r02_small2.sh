set -u
for o in "" "--opt resident_filter=2" "--opt early_weights=0" "--opt resident_filter=2 --opt early_weights=0"; do
  timeout 300 python bench.py --no-cpu-baseline --batch 64 $o > gpurun_out/r02_b64_tmp.json 2> gpurun_out/r02_b64_tmp.err; echo "b64 [$o] rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/r02_b64_tmp.json').read().strip().splitlines()[-1]);print(round(d['ms_per_step'],4), round(d['value']))")"
done
timeout 300 python bench.py --no-cpu-baseline --batch 64 --layer-report gpurun_out/r02_layers_b64_new.json > /dev/null 2>&1
