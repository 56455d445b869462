# final multi-GPU pass: bash tools/gpurun_scripts/r02_multigpu_final.sh N  (one gpurun --gpus N call): ResNet-50 weak + strong, default steps
set -u
N=$1
mkdir -p gpurun_out
run() {
  local name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r02f_mg_${name}_${N}gpu.json 2> gpurun_out/r02f_mg_${name}_${N}gpu.err
  echo "$name x$N rc=$? $(cut -c1-150 gpurun_out/r02f_mg_${name}_${N}gpu.json)"
}
run resnet50_weak
run resnet50_strong --scaling strong
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r02f_mg_reference_${N}gpu.json 2> gpurun_out/r02f_mg_reference_${N}gpu.err; echo "reference arm x$N rc=$? $(cut -c1-120 gpurun_out/r02f_mg_reference_${N}gpu.json)"
