#!/bin/bash
# Exchange-kernel / chunk-plan / wire-format sweep on N GPUs of one box (profiles/r02h_scaling.md):
#     gpurun --gpus N -- bash tools/exchange_sweep.sh N "<wire> <push_blocks> <gather_chunks> <push_tile>" ...
# wire 0 = xyz | label on the wire (16 B per point), 1 = t | label | ray index (12 B, points rebuilt on arrival).
N=$1
shift
for cfg in "$@"; do set -- $cfg
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --extra none --wire $1 --push-blocks $2 --gather-chunks $3 --push-tile ${4:-16384} 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('N $N wire $1 blocks $2 chunks $3 tile ${4:-16384}', d['value'], d['ms_per_step'], d['gather_bit_identical'], d['exchange']['nvlink_ingest_gbs_over_exchange_kernels'], d['roofline']['kernel_ms'], d['roofline'].get('compact_plus_exchange_ms'), d['e2e']['ms_per_step'])"
done
