"""ORACLE shim (test infrastructure only): minimal restatement of the
`compressai==1.2.4` surface the reference imports (pyproject.toml:16).
CompressAI is not installable in the build container; this follows the
published algorithm.  PARITY UNPINNED against the real package."""
__version__ = "1.2.4+oracle"
_entropy_coder = "ans"


def get_entropy_coder():
    return _entropy_coder


def available_entropy_coders():
    return ["ans"]
