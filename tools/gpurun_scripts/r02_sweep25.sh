set -u
run() { echo "== $*"; timeout 120 python tools/run_layers.py --network resnet50 --iters 5 "$@" 2>&1 | cut -c1-60,130-175,235-420; }
L=l4.1.conv3,l3.0.downsample,l3.0.conv1,l2.0.conv1,l2.1.conv1
run --layers $L
run --layers $L --opt warp_store=1
run --layers $L --opt warp_store=1 --opt stage_bufs=1
run --layers $L --opt n_stationary=1
