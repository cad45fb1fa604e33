timeout 200 python tools/layer_timeline.py 2>&1 | grep -v "cta"
timeout 120 python tools/step_parts.py 2>&1 | tail -6
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -3
