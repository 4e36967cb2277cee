#!/usr/bin/env python
"""Compress (and optionally decompress) a directory of PNGs with the B200-native CRDR path.

Same command line, files and columns as the reference's scripts/compress.py (:36-44, :96-139):
    python scripts/compress.py --config_path config/crdr.yaml --model_path crdr.pth.tar --img_dir IMG --save_dir OUT \
        -q 0.0 -b 3.84 --decompress -d cuda:0
writes OUT/<name>.bin, OUT/<name>.png, OUT/_bitrates.csv, OUT/_avg_bitrate.json.

Additions: --batch N codes up to N same-sized images per launch; under torchrun (RANK / WORLD_SIZE) the sorted
image list is sharded round-robin over the ranks (one process per GPU, no collective on the data path) and rank 0
merges the per-rank rows.
"""
import argparse
import json
import os
import sys
from glob import glob

import numpy as np
import pandas as pd
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import src  # noqa: E402,F401  (registers the model classes)
from src.models import build_comp_model  # noqa: E402
from src.utils import img_utils  # noqa: E402
from src.utils.codec_utils import load_byte_strings, save_byte_strings  # noqa: E402
from src.utils.logger import get_root_logger  # noqa: E402
from src.utils.options import TestConfig  # noqa: E402
from crdr_b200 import sharding  # noqa: E402


class CustomConfig(TestConfig):
    @classmethod
    def get_opt(cls):
        args = cls.arg_parse()
        cfg, text, _ = cls._file2dict_yaml(args["config_path"])
        merged = cls._merge_a_into_b(args, cfg)
        merged["is_train"] = False
        return cls(merged, cfg_text=text, filename=args["config_path"])

    @staticmethod
    def arg_parse():
        p = argparse.ArgumentParser()
        p.add_argument("--config_path", type=str, help="path to .yaml")
        p.add_argument("--model_path", type=str, help="path to model (.pth)")
        p.add_argument("--img_dir", type=str)
        p.add_argument("--save_dir", type=str)
        p.add_argument("-q", "--quality", type=float)
        p.add_argument("-b", "--beta", type=float)
        p.add_argument("--decompress", action="store_true")
        p.add_argument("-d", "--device", type=str, default="cuda:0")
        p.add_argument("--batch", type=int, default=1)
        return vars(p.parse_args())


def load_image(path):
    from PIL import Image
    arr = np.asarray(Image.open(path).convert("RGB"), dtype=np.float32) / 255.0  # ToTensor
    return torch.from_numpy(arr).permute(2, 0, 1).sub_(0.5).div_(0.5).unsqueeze(0)  # Normalize(.5, .5)


def main():
    opt = CustomConfig.get_opt()
    logger = get_root_logger()
    rank, world = sharding.rank_world()
    if world > 1 and str(opt.device).startswith("cuda"):
        opt.device = f"cuda:{int(os.environ.get('LOCAL_RANK', rank))}"
    os.makedirs(opt.save_dir, exist_ok=True)
    paths = sorted(glob(os.path.join(opt.img_dir, "*.png")))
    from PIL import Image
    sizes = []
    for p in paths:                     # header read only: balances the ranks by padded area and keeps equal shapes adjacent
        with Image.open(p) as im:
            sizes.append((im.height, im.width))
    mine = sharding.shard_balanced(paths, sizes, rank, world)

    model = build_comp_model(opt)
    model.load_learned_weight(ckpt_path=opt.model_path)
    model.codec_setup()

    # -q / -b: the reference compares them with 0.0 unconditionally (a missing flag is a TypeError there); here a missing
    # or negative -q is an error for the variable-rate models (their compress() needs rate_ind) and ignored by the
    # single-rate one, a missing or negative -b means the model's default beta
    uses_rate = getattr(model, "uses_rate", True)
    if uses_rate and (opt.quality is None or opt.quality < 0.0):
        raise SystemExit("compress.py: -q / --quality (0.0 .. rate_level - 1) is required for this model")
    beta_kw = {"beta": opt.beta} if (getattr(model, "uses_beta", True) and opt.beta is not None and opt.beta >= 0.0) else {}
    if world > 1 and os.environ.get("MASTER_ADDR") and not torch.distributed.is_initialized():
        torch.distributed.init_process_group("gloo")   # only for the final gather of the per-image rows

    rows = []
    # PNG decode ahead of the device (a few images in flight) and PNG encode behind it on host threads:
    # neither PIL's decoder nor cv2.imwrite holds the GIL while it works
    from collections import deque
    from concurrent.futures import ThreadPoolExecutor
    io_threads = max(2, min(8, (os.cpu_count() or 4) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world)))))
    pool = ThreadPoolExecutor(max_workers=io_threads)
    pending_writes = []

    def prefetched(paths_, depth):
        q = deque()
        it = iter(paths_)
        for p_ in it:
            q.append((p_, pool.submit(load_image, p_)))
            if len(q) >= depth:
                break
        while q:
            p_, fut = q.popleft()
            nxt = next(it, None)
            if nxt is not None:
                q.append((nxt, pool.submit(load_image, nxt)))
            yield p_, fut.result()

    images = prefetched(mine, max(2, 2 * int(opt.batch)))
    for group in sharding.same_shape_batches(images, opt.batch):
        x = torch.cat([im for _, im in group], dim=0)
        _, _, H, W = x.shape
        kwargs = {"rate_ind": opt.quality} if uses_rate else {}
        outs = model.compress_batch(x, **kwargs)
        bins = []
        for (path, _), out in zip(group, outs):
            name = os.path.basename(path)
            bin_path = os.path.join(opt.save_dir, name.replace(".png", ".bin"))
            save_byte_strings(bin_path, out["string_list"])
            nbytes = os.path.getsize(bin_path)
            sl = out["string_list"]
            rows.append({
                "img_name": name, "header_bit": len(sl[0]) * 8, "z_bit": len(sl[1]) * 8, "y_bit": len(sl[2]) * 8,
                "real_bit": nbytes * 8, "real_bpp": nbytes * 8 / H / W, "pred_z_bit": out["pred_z_bit"],
                "pred_y_bit": out["pred_y_bit"], "pred_bit": out["pred_z_bit"] + out["pred_y_bit"],
                "pred_bpp": out["pred_z_bpp"] + out["pred_y_bpp"], "num_pixel": H * W})
            bins.append(bin_path)
        if opt.decompress:
            imgs, _, _ = model.decompress_batch([load_byte_strings(b) for b in bins], **beta_kw)
            host = imgs.cpu()
            for (path, _), k in zip(group, range(len(group))):
                pending_writes.append(pool.submit(img_utils.imwrite, os.path.join(opt.save_dir, os.path.basename(path)), host[k:k + 1]))
            pending_writes = [f for f in pending_writes if not f.done() or f.result() is not None]
    for f in pending_writes:
        f.result()
    pool.shutdown()

    rows = sharding.gather_rows(rows, rank, world, opt.save_dir)
    if rank == 0:
        df = pd.json_normalize(sorted(rows, key=lambda r: r["img_name"]))
        df.to_csv(os.path.join(opt.save_dir, "_bitrates.csv"))
        avg = float(df["real_bpp"].mean()) if len(df) else float("nan")
        with open(os.path.join(opt.save_dir, "_avg_bitrate.json"), "w") as f:
            json.dump({"avg_bpp": avg}, f)
        logger.info(f"quality: {opt.quality}, beta: {opt.beta}")
        logger.info(f"num_image: {len(paths)}")
        logger.info(f"avg_bpp: {avg:.4f} [bpp]")
    if world > 1 and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
