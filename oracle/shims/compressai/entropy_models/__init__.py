from .entropy_models import EntropyBottleneck, EntropyModel, GaussianConditional

__all__ = ["EntropyModel", "EntropyBottleneck", "GaussianConditional"]
