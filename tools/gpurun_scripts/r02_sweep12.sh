set -u
run() { echo "== $*"; timeout 120 python tools/run_layers.py --network resnet50 --iters 5 "$@" 2>&1 | cut -c1-60,130-175,235-420; }
L=l3.1.conv2,l3.0.conv2,l4.1.conv2
run --layers $L
run --layers $L --opt keep_window=1
run --layers $L --opt keep_window=1 --opt cta_pairs=0
run --layers $L --opt keep_window=1 --opt paired_tiles=1 --opt cta_pairs=0
run --layers $L --opt max_bn=128
run --layers $L --opt max_bn=128 --opt keep_window=1
