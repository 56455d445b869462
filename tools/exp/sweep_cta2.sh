# knob sweep for the CTA-pair layers (development aid; run under gpurun)
L=l2.1.conv2,l3.1.conv2,l4.1.conv2,l3.1.conv1,l3.0.conv2
run() { echo "== $*"; env "$@" python tools/run_layers.py --layers $L --iters 3 2>&1 | cut -c1-60,150-330; }
run X=1
run LBC_TPS_KB=16
run LBC_TPS_KB=24
run LBC_TPS_KB=96
run LBC_MAX_WIN=2
run LBC_MAX_WIN=3
run LBC_MAX_STAGES=2
run LBC_MAX_STAGES=4
run LBC_STAGE_BUFS=2
run LBC_STAGE_BUFS=1
echo "== vgg"
python tools/run_layers.py --network vgg16 --layers conv2_2,conv3_2,conv4_2,conv5_2 --iters 3 2>&1 | cut -c1-60,150-330
LBC_TPS_KB=96 python tools/run_layers.py --network vgg16 --layers conv2_2,conv3_2,conv4_2,conv5_2 --iters 3 2>&1 | cut -c1-60,150-330
LBC_STAGE_BUFS=2 python tools/run_layers.py --network vgg16 --layers conv2_2,conv3_2,conv4_2,conv5_2 --iters 3 2>&1 | cut -c1-60,150-330
