from .engine import FSDPViT, FsdpUnit  # noqa: F401
from .layout import UnitLayout  # noqa: F401
from .optim import ShardedAdamW  # noqa: F401
from .graph import GraphedTrainStep  # noqa: F401
