set -u
run() { echo "== $*"; timeout 120 python tools/run_layers.py --iters 5 "$@" 2>&1 | cut -c1-60,130-175,235-420; }
N=conv1,l1.0.conv1,l1.0.conv2,l1.1.conv1
run --network resnet50 --layers $N
run --network resnet50 --layers $N --opt fold_bias=1
run --network vgg16 --layers conv1_1,conv1_2
run --network vgg16 --layers conv1_1,conv1_2 --opt fold_bias=1
