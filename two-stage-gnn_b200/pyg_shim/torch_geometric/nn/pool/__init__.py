from . import topk_pool  # noqa: F401
from .topk_pool import filter_adj, topk  # noqa: F401
