from crdr_b200.config import BaseConfig, ConfigDict, TestConfig  # noqa: F401
