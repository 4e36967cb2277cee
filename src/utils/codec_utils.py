from crdr_b200.codec_utils import *  # noqa: F401,F403
from crdr_b200.codec_utils import HeaderHandler, MultiRateHeaderHandler, load_byte_strings, save_byte_strings  # noqa: F401
