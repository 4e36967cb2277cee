"""Pins the portable oracle (oracle/crdr_oracle.py) to the unmodified reference modules.

Runs only where /root/reference exists (the build container); the GPU box relies on the committed goldens.
"""
import os
import sys

import pytest
import torch

from conftest import REFERENCE, ROOT

pytestmark = pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref_model():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
    added = REFERENCE not in sys.path
    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE)
    try:
        import src  # noqa: F401  (the reference's package)
        from src.models import build_comp_model
        from src.utils.options import BaseConfig
        cfg, _, _ = BaseConfig._file2dict_yaml(os.path.join(REFERENCE, "config", "crdr.yaml"))
        cfg["device"], cfg["is_train"] = "cpu", False
        torch.manual_seed(1234)
        m = build_comp_model(BaseConfig(cfg)).eval()
        m.codec_setup()
        yield m
    finally:
        sys.path.remove(REFERENCE)
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_compress_decompress_bit_equal(oracle, ref_model):
    sd = {k: v.detach().clone() for k, v in ref_model.state_dict().items()}
    x = torch.rand(1, 3, 72, 100, generator=torch.Generator().manual_seed(7)) * 2 - 1
    for q, beta in ((0.0, 3.84), (2.5, 0.0)):
        ref = ref_model.compress(x, rate_ind=q)
        mine = oracle.compress(sd, x, q)
        assert mine["string_list"] == ref["string_list"]
        assert torch.equal(mine["y_hat"], ref["y_hat"]) and torch.equal(mine["z_hat"], ref["z_hat"])
        assert torch.equal(mine["y_lik"], ref["y_likelihood"]) and torch.equal(mine["z_lik"], ref["z_likelihood"])
        assert abs(mine["pred_y_bit"] - ref["pred_y_bit"]) <= 1e-3 * abs(ref["pred_y_bit"])
        img_r, z_r, y_r = ref_model.decompress(ref["string_list"], beta=beta)
        img_o, z_o, y_o, _ = oracle.decompress(sd, mine["string_list"], beta)
        assert torch.equal(img_o, img_r) and torch.equal(y_o, y_r) and torch.equal(z_o, z_r)
        assert torch.equal(y_o, mine["y_hat"])  # the invariant of scripts/compress.py:126


def test_forward_eval_equal(oracle, ref_model):
    sd = {k: v.detach().clone() for k, v in ref_model.state_dict().items()}
    x = torch.rand(1, 3, 64, 128, generator=torch.Generator().manual_seed(3)) * 2 - 1
    with torch.no_grad():
        ref = ref_model.forward(x, 1.5, 2.56, is_train=False)
    eb, gc = oracle.entropy_models(sd)
    a = oracle.analysis(sd, x, 1.5, eb, gc)
    fake = oracle.g_s(sd, a["y_hat"], 1.5, 2.56).clamp(-1, 1)
    assert torch.equal(a["y"], ref["latent_code"]["y"]) and torch.equal(a["z"], ref["latent_code"]["z"])
    assert torch.equal(a["y_hat"], ref["quantized_code"]["y"])
    assert torch.equal(a["y_lik"], ref["q_likelihoods"]["y"])
    assert torch.equal(fake, ref["fake_images"])


def test_forward_train_equal_with_replayed_noise(oracle, ref_model):
    """Training-mode forward: the oracle, fed with the very noise values the reference draws under a seed
    (z first, in CompressAI's C x 1 x (N H W) order, then one draw per slice), reproduces the reference's dict."""
    sd = {k: v.detach().clone() for k, v in ref_model.state_dict().items()}
    n, h, w = 2, 64, 128
    x = torch.rand(n, 3, h, w, generator=torch.Generator().manual_seed(13)) * 2 - 1
    q, beta = 2.25, 1.28
    torch.manual_seed(77)
    with torch.no_grad():
        ref = ref_model.forward(x, q, beta, is_train=True)
    torch.manual_seed(77)
    zc, yc, hz, wz, hy, wy = 192, 320, h // 64, w // 64, h // 16, w // 16
    nz = torch.empty(zc, 1, n * hz * wz).uniform_(-0.5, 0.5).reshape(zc, n, hz, wz).permute(1, 0, 2, 3).contiguous()
    ny = torch.cat([torch.empty(n, yc // 10, hy, wy).uniform_(-0.5, 0.5) for _ in range(10)], dim=1)
    eb, gc = oracle.entropy_models(sd)
    mine = oracle.forward_train(sd, x, q, beta, {"z": nz, "y": ny}, eb, gc)
    for grp in ("likelihoods", "latent_code", "quantized_code", "q_likelihoods"):
        for k in ("y", "z"):
            assert torch.equal(mine[grp][k], ref[grp][k]), (grp, k)
    assert torch.equal(mine["fake_images"], ref["fake_images"])
    assert float(ref["fake_images"].abs().max()) > 0 and not torch.equal(ref["likelihoods"]["y"], ref["q_likelihoods"]["y"])


def test_discriminator_equals_reference_and_same_layout(oracle, ref_model):
    """ModuleListDiscriminator of CLIC21GVAEDiscriminators (config/crdr_stage_3.yaml:13-21).  The oracle's
    functional forward equals the unmodified reference module bit for bit, and this package's parameter tree has the
    reference's state_dict keys and shapes."""
    from src.models.discriminator import build_discriminator as build_ref     # ref_model's fixture put the reference on sys.path
    opt = dict(type="ModuleListDiscriminator", _subd_type="CLIC21GVAEDiscriminator", _num_subd=5, in_ch=3, out_ch=1, main_ch=64,
               norm_type="none")
    torch.manual_seed(99)
    ref = build_ref(dict(opt)).eval()
    x = torch.rand(2, 3, 64, 96, generator=torch.Generator().manual_seed(5)) * 2 - 1
    for q in (0.0, 2.0, 3.7):
        with torch.no_grad():
            want = ref(x, rate_ind=q)
        sub = {k[len(f"subD_list.{int(q)}."):]: v for k, v in ref.state_dict().items() if k.startswith(f"subD_list.{int(q)}.")}
        assert torch.equal(oracle.discriminator(sub, x), want)
    from crdr_b200.discriminator import build_discriminator as build_mine
    mine = build_mine(dict(opt))
    assert [(k, tuple(v.shape)) for k, v in mine.state_dict().items()] == [(k, tuple(v.shape)) for k, v in ref.state_dict().items()]


def test_training_gradients_equal_reference_autograd(oracle, ref_model):
    """The gradients the GPU training step is checked against (tests/test_gpu_train.py) are autograd through the oracle's
    forward_train; here they are pinned to autograd through the UNMODIFIED reference model (same seed -> same noise):
    rate + MSE loss, every parameter's gradient."""
    sd = {k: v.detach().clone() for k, v in ref_model.state_dict().items()}
    n, h, w = 1, 64, 64
    x = torch.rand(n, 3, h, w, generator=torch.Generator().manual_seed(14)) * 2 - 1
    q, beta = 1.5, 2.56

    def loss_of(out):
        bits = lambda lik: (-torch.log2(lik)).sum((1, 2, 3))
        bpp = (bits(out["likelihoods"]["y"]) + bits(out["likelihoods"]["z"])) / (h * w)
        return 0.8 * bpp.mean() + 150.0 * torch.mean(((x + 1) / 2 - (out["fake_images"] + 1) / 2) ** 2)

    ref_model.zero_grad()
    torch.manual_seed(78)
    loss_of(ref_model.forward(x, q, beta, is_train=True)).backward()
    ref = {k: p.grad.detach().clone() for k, p in ref_model.named_parameters() if p.grad is not None}
    ref_model.zero_grad()

    torch.manual_seed(78)
    zc, yc, hz, wz, hy, wy = 192, 320, h // 64, w // 64, h // 16, w // 16
    nz = torch.empty(zc, 1, n * hz * wz).uniform_(-0.5, 0.5).reshape(zc, n, hz, wz).permute(1, 0, 2, 3).contiguous()
    ny = torch.cat([torch.empty(n, yc // 10, hy, wy).uniform_(-0.5, 0.5) for _ in range(10)], dim=1)
    sdr = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    eb, gc = oracle.entropy_models(sdr)
    for p in eb.parameters():
        p.requires_grad_(True)
    with torch.enable_grad():
        loss_of(oracle.forward_train.__wrapped__(sdr, x, q, beta, {"z": nz, "y": ny}, eb, gc)).backward()
    mine = {k: v.grad for k, v in sdr.items() if v.is_floating_point() and v.grad is not None}
    for k, p in eb.named_parameters():
        if p.grad is not None:
            mine["entropy_model_z." + k] = p.grad
    checked = 0
    for k, g in ref.items():
        if k.endswith("quantiles"):
            continue          # aux parameter: no gradient from the main loss in either implementation
        assert k in mine, k
        den = float(g.abs().max())
        assert float((mine[k] - g).abs().max()) <= 1e-5 * max(den, 1e-12) + 1e-12, (k, float((mine[k] - g).abs().max()), den)
        checked += 1
    assert checked > 580


@pytest.mark.parametrize("stage", [1, 2])
def test_stage_models_bit_equal_and_same_layout(oracle, stage):
    """HyperpriorCharmModel (crdr_stage_1.yaml) and InterpCaHyperpriorCharmModel (crdr_stage_2.yaml).
    The oracle follows the unmodified reference bit for bit, and this package's parameter tree has the reference's
    state_dict keys, shapes and dtypes in the reference's order."""
    import fixtures
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE)
    try:
        import src  # noqa: F401  (the reference's package)
        from src.models import build_comp_model
        from src.utils.options import BaseConfig
        cfg, _, _ = BaseConfig._file2dict_yaml(os.path.join(REFERENCE, "config", f"crdr_stage_{stage}.yaml"))
        cfg["device"], cfg["is_train"] = "cpu", False
        torch.manual_seed(4321)
        ref = build_comp_model(BaseConfig(cfg)).eval()
        with torch.no_grad():
            fixtures.calibrate_(dict(ref.state_dict()), 0)
        ref.codec_setup()
        sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
        x = torch.rand(1, 3, 70, 96, generator=torch.Generator().manual_seed(9)) * 2 - 1
        q = None if stage == 1 else 2.5
        out = ref.compress(x) if stage == 1 else ref.compress(x, rate_ind=q)
        mine = oracle.compress(sd, x, q)
        assert mine["string_list"] == out["string_list"]
        assert torch.equal(mine["y_hat"], out["y_hat"]) and torch.equal(mine["y_lik"], out["y_likelihood"])
        img_r, z_r, y_r = ref.decompress(out["string_list"])
        img_o, z_o, y_o, _ = oracle.decompress(sd, mine["string_list"], None)
        assert torch.equal(img_o, img_r) and torch.equal(y_o, y_r) and torch.equal(z_o, z_r)
    finally:
        sys.path.remove(REFERENCE)
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    from crdr_b200.model import build_comp_model as build_mine
    m = build_mine(fixtures.crdr_opt("cuda:0", f"crdr_stage_{stage}.yaml"))
    with torch.no_grad():
        fixtures.calibrate_(dict(m.state_dict()), 0)   # same quantiles as the reference instance -> same table shapes
    m.codec_setup()
    got = {k: (tuple(v.shape), v.dtype) for k, v in m.state_dict().items()}
    want = {k: (tuple(v.shape), v.dtype) for k, v in sd.items()}
    assert list(got) == list(want) and got == want
