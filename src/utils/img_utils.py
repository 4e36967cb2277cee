from crdr_b200.img_utils import *  # noqa: F401,F403
from crdr_b200.img_utils import calc_psnr, imwrite, torch2npimg  # noqa: F401
