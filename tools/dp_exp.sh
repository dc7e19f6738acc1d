run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 30 --warmup 5 --no-score $EXTRA > gpurun_out/dp_$tag.json 2> gpurun_out/dp_$tag.err; python -c "
import json; d=json.load(open('gpurun_out/dp_$tag.json')); print('$tag', round(d['ms_per_step'],4), round(d['value']))"; }
EXTRA=""; run base A=1
EXTRA=""; run ctas4 NCCL_MAX_CTAS=4
EXTRA=""; run ctas8 NCCL_MAX_CTAS=8
EXTRA="--metrics-tier loss_only"; run lossonly A=1
