from crdr_b200.logger import *  # noqa: F401,F403
from crdr_b200.logger import get_root_logger, log_dict_items  # noqa: F401
