# every planner default against its alternative, whole ResNet-50 step (bench.py --opt ...), one box
set -u
mkdir -p gpurun_out
for o in "" "stage_bufs=2" "stage_bufs=1" "max_stages=6" "tps_kb=32" "tps_kb=64" "two_mma_warps=0" "n_stationary=0" "fold_bias=0" "small_teams=0" "tiles_per_iter2=0" "reverse=0" "four_acc=0" "max_win_stages=8" "resident_kb=64" "resident_kb=96" "warp_store=0" "cta_pairs=0" "tail_split=0" "early_weights=0" "pdl=0" ""; do
  args=""; for kv in $o; do args="$args --opt $kv"; done
  timeout 300 python bench.py --no-cpu-baseline $args > gpurun_out/r02_gs_tmp.json 2> gpurun_out/r02_gs_tmp.err
  echo "[$o] rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/r02_gs_tmp.json').read().strip().splitlines()[-1]);print(round(d['ms_per_step'],4), round(d['value']), d['parity'])" 2>/dev/null || tail -1 gpurun_out/r02_gs_tmp.err)"
done
