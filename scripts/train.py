#!/usr/bin/env python
"""Train a CRDR model with the B200-native step: the reference's scripts/train.py (:16-28) -- config -> build_trainer(opt) ->
trainer.train_loop() -- with the trainer classes of crdr_b200.trainers behind the reference's registry names.

    python scripts/train.py config/crdr_stage_2.yaml -d cuda:0 --total_iter 1000 [--img_dir DIR]
    torchrun --nproc-per-node 8 scripts/train.py config/crdr_stage_2.yaml          # data parallel, one process per GPU

Data: random 256 x 256 crops of the PNGs under --img_dir, or synthetic crops without it (the reference's OpenImages
pipeline, wandb and checkpoint rotation are outside the hot path this repository covers)."""
import argparse
import os
import sys
from glob import glob

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import src  # noqa: E402,F401  (registers the model / trainer classes)
from crdr_b200.config import BaseConfig  # noqa: E402
from crdr_b200.trainers import build_trainer  # noqa: E402


def load_image(path):
    from PIL import Image
    arr = np.asarray(Image.open(path).convert("RGB"), dtype=np.float32) / 255.0  # ToTensor
    return torch.from_numpy(arr).permute(2, 0, 1).sub_(0.5).div_(0.5).unsqueeze(0)  # Normalize(.5, .5)


def crops(img_dir, batch, size, seed):
    files = sorted(glob(os.path.join(img_dir, "*.png")))
    if not files:
        raise SystemExit(f"no PNG files under {img_dir}")
    rng = np.random.default_rng(seed)
    cache = {}
    while True:
        out = []
        for _ in range(batch):
            f = files[rng.integers(len(files))]
            if f not in cache:
                cache[f] = load_image(f)
            img = cache[f]                                   # [1, 3, H, W] in [-1, 1]
            h, w = img.shape[-2:]
            y, x = rng.integers(0, h - size + 1), rng.integers(0, w - size + 1)
            out.append(img[0, :, y:y + size, x:x + size])
        yield {"real_images": torch.stack(out)}


def main():
    p = argparse.ArgumentParser()
    p.add_argument("config_path")
    p.add_argument("-d", "--device", default=None)
    p.add_argument("--total_iter", type=int, default=None)
    p.add_argument("--img_dir", default=None)
    p.add_argument("--batch_size", type=int, default=8)
    p.add_argument("--patch_size", type=int, default=256)
    p.add_argument("--exp", default=None)
    a = p.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    device = a.device or f"cuda:{local}"
    torch.cuda.set_device(torch.device(device))
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device(device))
    opt = BaseConfig.fromfile(a.config_path, device=device, is_train=True)
    opt["batch_size"], opt["patch_size"] = a.batch_size, a.patch_size
    opt["exp"] = a.exp or os.path.basename(a.config_path).split(".")[0]
    if a.total_iter is not None:
        opt["total_iter"] = a.total_iter
    trainer = build_trainer(opt)
    batches = crops(a.img_dir, a.batch_size, a.patch_size, seed=rank) if a.img_dir else None
    trainer.train_loop(batches)
    if rank == 0:
        print("saved", trainer.save(int(opt["total_iter"])))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
