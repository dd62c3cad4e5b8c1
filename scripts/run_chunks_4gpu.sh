run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 4 --steps 60 --no-extras --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$2', d['value'], d['ms_per_step'])"; }
run 29601 K1
SHB_EXCHANGE_CHUNKS=2 run 29602 K2
SHB_EXCHANGE_CHUNKS=4 run 29603 K4
SHB_PRE_OVERLAP=4 run 29604 PRE4
SHB_EXCHANGE_CHUNKS=2 SHB_OVL_PRIO=1 run 29605 K2prio
