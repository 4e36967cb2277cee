from crdr_b200.trainers import build_trainer  # noqa: F401
