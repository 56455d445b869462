set -u
run() { echo "== $*"; timeout 120 python tools/run_layers.py --network resnet50 --iters 5 "$@" 2>&1 | cut -c1-60,130-175,235-420; }
L=l1.1.conv3,l2.1.conv3,l3.1.conv3,l4.1.conv3,l3.0.downsample
run --layers $L
run --layers $L --opt epi_split=1
