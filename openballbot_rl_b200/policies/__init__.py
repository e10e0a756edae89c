"""Policy network architectures (ballbot_rl/policies/__init__.py:1-11): registers the feature extractor as policy plugin "mlp"."""
from ..core.registry import ComponentRegistry
from .mlp_policy import Extractor

if "mlp" not in ComponentRegistry.list_policies():
    ComponentRegistry.register_policy("mlp", Extractor)

__all__ = ["Extractor"]
