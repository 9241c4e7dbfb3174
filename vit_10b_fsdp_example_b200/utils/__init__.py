from .meters import SmoothedValue  # noqa: F401
from .schedule import WarmupCosineSchedule, get_warmup_cosine_scheduler  # noqa: F401
