"""Small synthetic class-incremental run through template.main (debug / demo): prints the per-epoch training stats."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import wacil_b200 as w

if __name__ == "__main__":
    argv = sys.argv[1:] or ["--data_set", "synthetic", "--nb_classes", "6", "--num_bases", "2", "--increment", "2",
                            "--backbone", "resnet20", "--batch_size", "32", "--num_epochs", "10", "--lr", "0.05",
                            "--memory_size", "60", "--synthetic_per_class", "128", "--synthetic_val_per_class", "32",
                            "--eval_every_epoch", "5"]
    print(w.template.cli(argv))
