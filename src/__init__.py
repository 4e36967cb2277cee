"""Drop-in alias of the reference's top-level ``src`` package: importing it registers the B200-native
classes under the reference's registry names, so ``scripts/compress.py`` (reference or ours) runs unchanged."""
import crdr_b200.model  # noqa: F401
import crdr_b200.discriminator  # noqa: F401
import crdr_b200.trainers  # noqa: F401
