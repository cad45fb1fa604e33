"""template.main on the synthetic data set (the configuration of tests/test_gpu_multi.py) on one GPU: prints the
per-task accuracies.  WACIL_ENGINE=layer|chain selects the engine."""
import argparse
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import wacil_b200 as w  # noqa: E402

p = argparse.ArgumentParser(parents=[w.get_args_parser()])
w.train.add_runner_args(p)
args = p.parse_args(["--data_set", "synthetic", "--nb_classes", "6", "--num_bases", "2", "--increment", "2",
                     "--backbone", "resnet20", "--batch_size", sys.argv[1] if len(sys.argv) > 1 else "32",
                     "--num_epochs", "10", "--lr", "0.05", "--memory_size", "40", "--synthetic_per_class", "96",
                     "--synthetic_val_per_class", "32", "--eval_every_epoch", "5", "--max_tasks", "2", "--checkpoint_dir",
                     tempfile.mkdtemp()])
print(os.environ.get("WACIL_ENGINE", "layer"), "batch", args.batch_size, "acc", w.template.main(args, log=lambda *a, **k: None))
