set -u
mkdir -p gpurun_out
for o in "" "--opt early_weights=2" "" "--opt early_weights=2" "--opt early_weights=0"; do
  timeout 300 python bench.py --no-cpu-baseline $o > gpurun_out/r02_b23_tmp.json 2> gpurun_out/r02_b23_tmp.err; echo "[$o] rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/r02_b23_tmp.json').read().strip().splitlines()[-1]);print(round(d['ms_per_step'],4), round(d['value']))")"
done
