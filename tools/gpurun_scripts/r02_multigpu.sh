# r02 multi-GPU pass: bash tools/gpurun_scripts/r02_multigpu.sh N  (one gpurun --gpus N call)
set -u
N=$1
mkdir -p gpurun_out
run() { # name, extra args
  local name=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r02_mg_${name}_${N}gpu.json 2> gpurun_out/r02_mg_${name}_${N}gpu.err
  echo "$name x$N rc=$? $(cut -c1-150 gpurun_out/r02_mg_${name}_${N}gpu.json)"
}
run resnet50_weak
run resnet50_strong --scaling strong
run resnet18_weak --network resnet18
run vgg16_weak --network vgg16
run mobilenet_v2_weak --network mobilenet_v2
