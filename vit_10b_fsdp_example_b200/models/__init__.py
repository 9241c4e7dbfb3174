from . import vit  # noqa: F401
