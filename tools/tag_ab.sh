cd a-pytorch*/csrc; for c in 18 19 20; do timeout 60 ./bin/test_tc $c time; done; cd ../..
timeout 120 python tools/step_parts.py 2>&1 | tail -6
timeout 400 python -m pytest tests/test_gpu_engine.py tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -3
