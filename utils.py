"""Drop-in module mirroring the reference's top-level ``utils.py`` API on the B200-native engine:
``get_warmup_cosine_scheduler`` (utils.py:11-21), ``save_ckpt`` / ``load_ckpt`` (utils.py:25-43),
``FakeImageNetDataset`` (utils.py:46-55), ``SmoothedValue`` (utils.py:60-102)."""
from vit_10b_fsdp_example_b200.data import FakeImageNetDataset  # noqa: F401
from vit_10b_fsdp_example_b200.utils.checkpoint import load_ckpt, save_ckpt  # noqa: F401
from vit_10b_fsdp_example_b200.utils.meters import SmoothedValue  # noqa: F401
from vit_10b_fsdp_example_b200.utils.schedule import get_warmup_cosine_scheduler  # noqa: F401
