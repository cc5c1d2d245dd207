python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
for cfg in "4 1" "4 3" "3 2" "6 3" "2 1" "2 3" "8 4"; do set -- $cfg
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --extra none --gather-chunks $1 --gather-taper $2 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('chunks $1 taper $2', d['value'], d['ms_per_step'], d['gather_bit_identical'])"
done
