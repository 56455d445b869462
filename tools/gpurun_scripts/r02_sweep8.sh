set -u
run() { echo "== $*"; timeout 120 python tools/run_layers.py --network resnet50 --iters 5 "$@" 2>&1 | cut -c1-60,130-175,235-420; }
L=l4.0.conv3,l3.0.conv1,l3.0.downsample,l4.0.conv1,l3.0.conv3
run --layers $L
run --layers $L --opt cta_pairs=1
run --layers $L --opt cta_pairs=0
N=conv1,l1.0.conv1,l1.0.conv2,l1.1.conv1
run --layers $N
run --layers $N --opt warp_store=0
