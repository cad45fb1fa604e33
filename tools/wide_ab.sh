for w in 1 0 1 0; do echo "== WACIL_WIDE=$w"; WACIL_WIDE=$w timeout 120 python tools/step_parts.py 2>&1 | tail -6; done
WACIL_WIDE=1 timeout 300 python -m pytest tests/test_gpu_engine.py -x -q -m gpu 2>&1 | tail -3
