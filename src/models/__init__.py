from crdr_b200.model import build_comp_model, build_subnet  # noqa: F401

__all__ = ["build_comp_model"]
