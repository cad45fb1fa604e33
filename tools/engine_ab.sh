for eng in layer chain; do
  echo "=== WACIL_ENGINE=$eng"
  WACIL_ENGINE=$eng python bench.py --steps 50 --warmup 10 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("bench", d["ms_per_step"], d["e2e"]["ms_per_step"], d["launches_per_step"])'
done
WACIL_ENGINE=layer python -m pytest tests -m gpu -q 2>&1 | tail -3
