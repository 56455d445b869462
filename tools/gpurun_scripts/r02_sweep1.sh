set -u
L=l1.1.conv3,l2.1.conv3,l3.1.conv3,l4.1.conv3,l3.0.downsample,l2.0.downsample
run() { echo "== $*"; timeout 120 python tools/run_layers.py --network resnet50 --layers $L --iters 5 "$@" 2>&1 | cut -c1-60,150-400; }
run --opt epi_split=0
run --opt epi_split=0 --opt n_stationary=0
run --opt epi_split=0 --opt max_bn=128
run --opt epi_split=0 --opt max_bn=128 --opt warp_store=1
run --opt epi_split=0 --opt warp_store=0
run --opt epi_split=1 --opt warp_store=0
run --opt epi_split=0 --opt max_bn=128 --opt fold_bias=0
run --opt epi_split=0 --opt max_bn=64
