from crdr_b200.discriminator import build_discriminator  # noqa: F401
