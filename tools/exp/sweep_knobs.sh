L=conv1,l1.1.conv2,l1.1.conv3,l2.1.conv2,l3.1.conv2
run() { echo "== $*"; env "$@" python tools/run_layers.py --layers $L --iters 3 2>&1 | cut -c1-200; }
run X=1
run LBC_TPS_KB=48
run LBC_TPS_KB=48 LBC_MAX_WIN=2 LBC_MAX_STAGES=2
run LBC_STAGE_BUFS=1
run LBC_TPS_KB=48 LBC_STAGE_BUFS=1 LBC_MAX_WIN=2 LBC_MAX_STAGES=2
run LBC_MAX_WIN=3
run LBC_TPS_KB=8
