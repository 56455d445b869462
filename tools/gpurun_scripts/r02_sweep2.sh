set -u
L=conv1,l1.0.conv1,l1.1.conv1,l1.1.conv2,l2.1.conv1
run() { echo "== $*"; timeout 120 python tools/run_layers.py --network resnet50 --layers $L --iters 5 "$@" 2>&1 | cut -c1-60,130-175,235-420; }
run
run --opt two_mma_warps=0
run --opt two_mma_warps=0 --opt stage_bufs=2
run --opt stage_bufs=2
run --opt two_mma_warps=0 --opt stage_bufs=1
run --opt two_mma_warps=0 --opt tiles_per_iter2=0
