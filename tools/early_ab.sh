run2() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 200 --warmup 10 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['replicas_identical'])"; }
echo "== N=2 early=1 blocks=32"; WACIL_OPT_EARLY=1 run2 29511
echo "== N=2 early=1 blocks=8"; WACIL_OPT_EARLY=1 WACIL_OPT_EARLY_BLOCKS=8 run2 29512
echo "== N=2 early=2 (full grid)"; WACIL_OPT_EARLY=2 run2 29513
echo "== N=2 early=0"; WACIL_OPT_EARLY=0 run2 29514
for e in 1 0; do echo "== N=1 early=$e"; WACIL_OPT_EARLY=$e timeout 300 python bench.py --steps 200 --warmup 10 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])"; done
