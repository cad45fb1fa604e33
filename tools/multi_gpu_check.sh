timeout 500 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 2>/dev/null | tail -1
