from crdr_b200.registry import *  # noqa: F401,F403
