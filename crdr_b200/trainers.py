"""Trainer plugins under the reference's registry names (src/trainer/*): ``build_trainer(opt)`` resolves ``opt.trainer.type``
through TRAINER_REGISTRY exactly like src/trainer/__init__.py:9-26, and the classes keep the reference's step interface

    trainer.optimize_parameters(current_iter, data_dict)   # data_dict = {"real_images": [N, 3, H, W] in [-1, 1]}
    trainer.train_loop(), trainer.save(current_iter)

with the step itself running on the CUDA engines (crdr_b200.train).  What the reference's BaseTrainer does around the step
(datasets, wandb, csv loggers, checkpoint rotation; base_trainer.py:18-215) is out of scope: the loop here
takes any iterable of batches and writes ``{'iter', 'comp_model'}`` checkpoints in the reference layout."""
import os
from copy import deepcopy

import torch

from .discriminator import build_discriminator
from .logger import get_root_logger
from .model import build_comp_model
from .registry import TRAINER_REGISTRY
from .train import CodecTrainer, GanCodecTrainer


def build_trainer(opt):
    """src/trainer/__init__.py:9-26."""
    if not opt.get("trainer"):
        raise ValueError('"trainer_type" key is not supported. Please use trainer.type')
    kw = dict(deepcopy(opt["trainer"]))
    return TRAINER_REGISTRY.get(kw.pop("type"))(opt, **kw)


def _loss_kwargs(opt):
    """Loss / optimiser settings in the reference's yaml schema (config/crdr_stage_2.yaml:14-33) -> CodecTrainer kwargs."""
    loss, optim = opt["loss"], opt["optim"]
    rate = loss["rate_loss"]
    if rate["type"] not in ("HificVariableRateLoss", "HificRateLoss"):
        raise NotImplementedError(f"rate loss {rate['type']} is not lowered")
    dist = loss["distortion_loss"]
    if dist["type"] != "MSELoss" or dist.get("mse_scale", "0_1") != "0_1":
        raise NotImplementedError("only MSELoss on the 0..1 scale is lowered")
    g = optim["g_optimizer"]
    if g["type"] != "Adam":
        raise NotImplementedError("only Adam is lowered")
    as_list = lambda v: list(v) if isinstance(v, (list, tuple)) else float(v)
    percep = loss.get("perceptual_loss")
    if percep and (percep["type"] != "LPIPSLoss" or percep.get("range_norm", False)):
        raise NotImplementedError("only LPIPSLoss on [-1, 1] inputs is mapped (to its weight-free stand-in)")
    return dict(lr=float(g["lr"]), clip_max_norm=optim.get("clip_max_norm"), lambda_mse=float(dist["loss_weight"]),
                perceptual_weight=float(percep["loss_weight"]) if percep else 0.0,
                rate_lambda_a=as_list(rate["lambda_A"]), rate_lambda_b=as_list(rate["lambda_B"]), target_rate=as_list(rate["target_rate"]),
                aux_lr=float(optim.get("aux_optimizer", {}).get("lr", 1e-3)))


class _TrainerBase:
    def __init__(self, opt):
        self.opt, self.device, self.logger = opt, opt["device"], get_root_logger()
        self.comp_model = build_comp_model(opt)
        if opt.get("pretrained_weight_path") and os.path.exists(opt["pretrained_weight_path"]):
            self.comp_model.load_learned_weight(opt["pretrained_weight_path"])
        if opt["loss"].get("perceptual_loss"):
            self.logger.warning("perceptual_loss: LPIPS needs pretrained AlexNet weights that are not available here; the weight-free "
                                "stand-in of the oracle stack (per-image mean squared difference, oracle/shims/lpips.py) is used")
        self.milestones = list(opt["optim"].get("g_scheduler", {}).get("milestones", []))
        self.gamma = float(opt["optim"].get("g_scheduler", {}).get("gamma", 0.1))
        self.base_lr = float(opt["optim"]["g_optimizer"]["lr"])

    def run_comp_model(self, data_dict):
        raise NotImplementedError("the forward is part of the lowered step (optimize_parameters)")

    def _schedule(self, current_iter):
        """MultiStepLR (optimizer/__init__.py, torch.optim.lr_scheduler.MultiStepLR)."""
        lr = self.base_lr * self.gamma ** sum(1 for m in self.milestones if current_iter >= m)
        if lr != self.step_impl.lr:
            self.step_impl.set_lr(lr)

    def optimize_parameters(self, current_iter, data_dict):
        x = data_dict["real_images"].to(self.device, torch.float32, non_blocking=True).contiguous()
        self._schedule(current_iter)
        return self.step_impl.train_step(x, q=data_dict.get("rate_ind"), **({"beta": data_dict["beta"]} if "beta" in data_dict else {}))

    def train_loop(self, batches=None):
        """batches: iterable of data_dicts (default: an endless stream of synthetic crops, for smoke runs)."""
        if batches is None:
            batches = self._synthetic()
        start, total = int(self.opt.get("start_iter", 0)), int(self.opt["total_iter"])
        log_step, save_step = int(self.opt.get("log_step", 100)), int(self.opt.get("save_step", 5000))
        for itr, data in zip(range(start + 1, total + 1), batches):
            loss = self.optimize_parameters(itr, data)
            if itr % log_step == 0:
                self.logger.info(f"iter {itr}: " + ", ".join(f"{k} {float(v):.4f}" for k, v in loss.items()))
            if itr % save_step == 0:
                self.save(itr)

    def validation(self, dataloader, max_sample_size=1 << 30, **kw):
        """base_trainer.py:162-169: evaluate the current parameters with the model's own validation() (evaluation-mode
        forward on the inference engines, one row of bpp / PSNR per image and quality level).  dataloader: a sequence of
        data_dicts {"real_images": [1, 3, H, W]}."""
        self.step_impl.sync_to_model()
        self.comp_model.invalidate_engine()
        self.comp_model.codec_setup()
        return self.comp_model.validation(dataloader, max_sample_size, **kw)

    def _synthetic(self):
        g = torch.Generator().manual_seed(0)
        n, s = int(self.opt.get("batch_size", 8)), int(self.opt.get("patch_size", 256))
        while True:
            yield {"real_images": torch.rand(n, 3, s, s, generator=g) * 2 - 1}

    def save(self, current_iter):
        self.step_impl.sync_to_model()
        root = os.path.join(self.opt.get("ckpt_root", "./checkpoint"), str(self.opt.get("exp", "exp")), "model")
        os.makedirs(root, exist_ok=True)
        path = os.path.join(root, f"comp_model_iter{current_iter}.pth.tar")
        torch.save({"iter": current_iter, "comp_model": self.comp_model.state_dict()}, path)
        if self.opt.get("keep_training_state", False):      # rate_distortion_trainer.py:104-117: optimiser state next to the model
            torch.save({"iter": current_iter, "training_state": self.step_impl.training_state()},
                       os.path.join(root, f"training_state_iter{current_iter}.pth.tar"))
        return path

    def load_checkpoint(self, exp, itr):
        """Resume (base_trainer.py:102-108): parameters, Adam moments, step count and loss scale of `training_state_iter{itr}`."""
        root = os.path.join(self.opt.get("ckpt_root", "./checkpoint"), str(exp), "model")
        state = torch.load(os.path.join(root, f"training_state_iter{itr}.pth.tar"), map_location="cpu")
        self.step_impl.load_training_state(state["training_state"])


@TRAINER_REGISTRY.register()
class RateDistortionTrainer(_TrainerBase):
    """rate_distortion_trainer.py:16-101 (stage 1 / stage 2)."""

    def __init__(self, opt):
        super().__init__(opt)
        self.step_impl = CodecTrainer(self.comp_model, device=self.device, **_loss_kwargs(opt))


@TRAINER_REGISTRY.register()
class MultirateBetaCondHrrGanRateDistortionTrainer(_TrainerBase):
    """multirate_hr_rgan_beta_cond_rate_distortion_trainer.py:10-114 (stage 3)."""

    def __init__(self, opt, relative_score_rate_delta=1):
        super().__init__(opt)
        self.discriminator = build_discriminator(dict(opt["discriminator"]))
        gan = opt["loss"]["gan_loss"]
        if gan["type"] != "VanillaGANLoss":
            raise NotImplementedError(f"GAN loss {gan['type']} is not lowered")
        self.step_impl = GanCodecTrainer(self.comp_model, self.discriminator, device=self.device, lambda_gan=float(gan["loss_weight"]),
                                         d_lr=float(opt["optim"]["d_optimizer"]["lr"]), relative_score_rate_delta=relative_score_rate_delta,
                                         **_loss_kwargs(opt))
