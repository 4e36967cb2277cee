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
