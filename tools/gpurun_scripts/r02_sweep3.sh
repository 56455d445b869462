set -u
L=conv1,l1.1.conv2
run() { echo "== $*"; timeout 120 python tools/run_layers.py --network resnet50 --layers $L --iters 5 "$@" 2>&1 | cut -c1-60,130-175,235-420; }
run
run --opt stage_bufs=1
run --opt stage_bufs=2
run --opt stage_bufs=3
run --opt max_win_stages=6
run --opt max_win_stages=8
run --opt stage_bufs=2 --opt max_win_stages=10
echo "== vgg / resnet18 window layers"
timeout 120 python tools/run_layers.py --network vgg16 --layers conv1_1,conv1_2,conv2_1,conv2_2 --iters 3 2>&1 | cut -c1-60,130-175,235-420
timeout 120 python tools/run_layers.py --network resnet18 --layers l1.0.conv1,l2.0.conv2,l3.0.conv2 --iters 3 2>&1 | cut -c1-60,130-175,235-420
