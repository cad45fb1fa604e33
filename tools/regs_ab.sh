cd a-pytorch-tutorial-to-class-incremental-learning_b200/csrc
cp libwacil.so libwacil_base.so
for R in 96 112 128; do
  cp libwacil_r$R.so libwacil.so
  cd ../..
  for cfg in "64 4" "64 8" "4 4" "2 2" "1 1"; do set -- $cfg
    echo "regs $R FWD=$1 BWD=$2: $(WACIL_CHAIN_FWD=$1 WACIL_CHAIN_BWD=$2 python bench.py --steps 50 --warmup 10 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["launches_per_step"])')"
  done
  cd a-pytorch-tutorial-to-class-incremental-learning_b200/csrc
done
cp libwacil_base.so libwacil.so
