from .meters import AverageMeter, ProgressMeter, accuracy, adjust_learning_rate  # noqa: F401
from .checkpoint import save_checkpoint, export_state_dict, load_checkpoint  # noqa: F401
