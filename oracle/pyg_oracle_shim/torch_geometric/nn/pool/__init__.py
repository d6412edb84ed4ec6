from . import topk_pool  # noqa: F401
