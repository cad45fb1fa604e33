"""B200-native WA class-incremental training step behind the reference's own API surface.

Reference: G-U-N/a-PyTorch-Tutorial-to-Class-Incremental-Learning (template.py / resnet.py / utils.py).
All arithmetic lives in csrc/libwacil.so (hand-written sm_100a CUDA: tcgen05/TMEM convolutions fed by TMA, fused
KD / CE+KD, weight-align, SGD, herding).  Importing the package on a host without the built library raises."""
from . import _lib
from ._lib import WacilError, build  # noqa: F401

_lib.lib()  # fail loudly at import time if the CUDA library is missing

from . import checkpoint, data, ops  # noqa: E402,F401
from .model import CilClassifier, CilModel, SoftTarget, freeze_parameters  # noqa: E402,F401
from .data import ClassIncrementalScenario, DeviceTaskSet, build_dataset  # noqa: E402,F401
from .rehearsal import (RehearsalMemory, herd_closest_to_barycenter, herd_closest_to_cluster,  # noqa: E402,F401
                        herd_random)
from .train import (TaskTrainer, accuracy, eval, evaluate, extract_features, get_args_parser,  # noqa: E402,F401
                    init_distributed_mode, init_seed, train_task, update_model)
from . import template  # noqa: E402,F401
