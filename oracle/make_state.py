"""ORACLE side -- test infrastructure only.  Seeded random-init checkpoint in the reference layout WITHOUT the product
package: bench.py's reference arm must not import or map anything of crdr_b200.

Shapes and key order come from tests/golden/state_dict_layout_crdr.json (dumped from the unmodified reference,
tests/golden/make_golden.py).  Values: PyTorch-default-like uniform(-1/sqrt(fan_in), 1/sqrt(fan_in)) for conv / linear
weights and biases, N(0, 0.02) for the decoder (weight_init: True, elic_interpca_beta_cond_autoencoder.py:30-40,147-148),
CompressAI's EntropyBottleneck initial values, then the same `calibrate_` as tests/fixtures.py so the entropy path
is exercised.  The timing of the CPU path does not depend on the exact values."""
import json
import math
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def random_state_dict(seed=0, calibrated=True, config="crdr.yaml"):
    layout = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_layout_crdr.json")))
    if config == "crdr_stage_2.yaml":
        # InterpCaHyperpriorCharmModel = the crdr.yaml model without the decoder's beta conditioning (MLP + 27 projections)
        layout = {k: v for k, v in layout.items() if not (k.startswith("decoder.mlp.") or ".proj_" in k)}
    elif config != "crdr.yaml":
        raise ValueError(config)
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, (shape, dtype) in layout.items():
        dt = getattr(torch, dtype)
        if key.startswith("entropy_model_") and not key.startswith("entropy_model_z._") and "quantiles" not in key:
            sd[key] = torch.zeros(shape, dtype=dt)          # tables / bounds: rebuilt by oracle.entropy_models()
            continue
        if not dt.is_floating_point:
            sd[key] = torch.zeros(shape, dtype=dt)
            continue
        if key.startswith("entropy_model_z._matrix"):
            i = int(key[-1])
            widths = (1, 3, 3, 3, 3, 1)
            scale = 10 ** (1 / 5)
            sd[key] = torch.full(shape, float(np.log(np.expm1(1 / scale / widths[i + 1]))), dtype=dt)
        elif key.startswith("entropy_model_z._bias"):
            sd[key] = torch.rand(shape, generator=g, dtype=dt) - 0.5
        elif key.startswith("entropy_model_z._factor"):
            sd[key] = torch.zeros(shape, dtype=dt)
        elif key == "entropy_model_z.quantiles":
            sd[key] = torch.tensor([-10.0, 0.0, 10.0]).repeat(shape[0], 1, 1)
        elif "interp_ca_list" in key:
            # InterpChAtt init (interp_channel_attention.py:17-37): softplus(weight) == 1, bias == 0
            sd[key] = torch.full(shape, float(np.log(np.e - 1)) if key.endswith("weight") else 0.0, dtype=dt)
        elif key.startswith("decoder.") and len(shape) >= 2:
            sd[key] = torch.randn(shape, generator=g, dtype=dt) * 0.02
        elif len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            b = 1.0 / math.sqrt(fan_in)
            sd[key] = (torch.rand(shape, generator=g, dtype=dt) * 2 - 1) * b
        else:
            sd[key] = (torch.rand(shape, generator=g, dtype=dt) * 2 - 1) * 0.05
    if calibrated:
        import sys
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import fixtures  # pure-torch helpers; imports nothing of the product at module level
        fixtures.calibrate_(sd, seed)
    return sd
