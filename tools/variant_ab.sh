cd a-pytorch-tutorial-to-class-incremental-learning_b200/csrc
cp libwacil.so libwacil_base.so
for V in A B C B A; do
  cp libwacil_v$V.so libwacil.so
  cd ../..
  echo "variant $V: $(WACIL_ENGINE=layer python bench.py --steps 100 --warmup 10 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["launches_per_step"], d["roofline"]["us"])')"
  cd a-pytorch-tutorial-to-class-incremental-learning_b200/csrc
done
cp libwacil_base.so libwacil.so
