"""Generates the committed golden fixtures.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

* rans_vectors.npz      -- symbols / indexes / expected bytes from the oracle coder (exact integer KATs)
* entropy_tables.npz    -- GaussianConditional and EntropyBottleneck CDF tables from the oracle restatement
* codec_<name>.npz      -- outputs of the UNMODIFIED REFERENCE modules (imported from /root/reference on top of the
                           oracle's CompressAI restatement) for a seeded checkpoint (tests/fixtures.build_model layout)
                           and seeded images: bitstreams, symbols, indexes, bit counts, reconstruction (uint8).
The checkpoint itself (511 MB) is not committed: it is regenerated from the seed on any machine.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle", "shims")]
REFERENCE = "/root/reference"


def rans_vectors():
    from compressai import ans
    from compressai.entropy_models import GaussianConditional
    from compressai.models import get_scale_table
    gc = GaussianConditional(None, scale_bound=0.11)
    gc.update_scale_table(get_scale_table(), force=True)
    rng = np.random.default_rng(1)
    n = 6000
    idx = rng.integers(0, 64, n).astype(np.int32)
    sym = np.rint(rng.normal(size=n) * get_scale_table().numpy()[idx] * 1.3).astype(np.int32)
    sym[::211] = 4000        # escapes above the table
    sym[7::307] = -90000     # escapes below the table
    sym[11] = 2 ** 20
    cdf, lens, offs = gc._quantized_cdf.numpy(), gc._cdf_length.numpy(), gc._offset.numpy()
    stream = ans.RansEncoder().encode_with_indexes(sym, idx, cdf, lens, offs)
    np.savez_compressed(os.path.join(HERE, "rans_vectors.npz"), symbols=sym, indexes=idx,
                        stream=np.frombuffer(stream, dtype=np.uint8))
    eb_rows = None
    from compressai.entropy_models import EntropyBottleneck
    torch.manual_seed(5)
    eb = EntropyBottleneck(8)
    with torch.no_grad():
        eb.quantiles[:, 0, 0] = -torch.arange(8).float() - 2.3
        eb.quantiles[:, 0, 2] = torch.arange(8).float() * 1.5 + 3.1
        eb.quantiles[:, 0, 1] = torch.linspace(-0.4, 0.4, 8)
    eb.update(force=True)
    np.savez_compressed(
        os.path.join(HERE, "entropy_tables.npz"),
        gc_cdf_sha256=np.frombuffer(hashlib.sha256(cdf.astype(np.int32).tobytes()).digest(), dtype=np.uint8),
        gc_cdf_rows=cdf[[0, 1, 17, 40, 63]], gc_lengths=lens, gc_offsets=offs,
        **{"ebp" + k: v.numpy() for k, v in eb.state_dict().items() if k.startswith(("_matrix", "_bias", "_factor", "quantiles"))},
        eb_cdf=eb._quantized_cdf.numpy(), eb_lengths=eb._cdf_length.numpy(), eb_offsets=eb._offset.numpy())


def reference_model(sd):
    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE)
    try:
        import src  # noqa: F401
        from src.models import build_comp_model
        from src.utils.options import BaseConfig
        cfg, _, _ = BaseConfig._file2dict_yaml(os.path.join(REFERENCE, "config", "crdr.yaml"))
        cfg["device"], cfg["is_train"] = "cpu", False
        m = build_comp_model(BaseConfig(cfg)).eval()
        m.load_state_dict(sd)
        m.codec_setup()
        return m
    finally:
        sys.path.remove(REFERENCE)
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def codec_goldens():
    import fixtures
    import crdr_oracle as orc
    cases = [("calib_96x160_q1.5_b2.56", True, 96, 160, 1.5, 2.56), ("calib_70x100_q0_b3.84", True, 70, 100, 0.0, 3.84),
             ("raw_64x64_q4_b0", False, 64, 64, 4.0, 0.0)]
    models = {}
    for name, calibrated, h, w, q, beta in cases:
        if calibrated not in models:
            _, sd = fixtures.build_model(seed=0, calibrated=calibrated, device="cuda:0")
            models[calibrated] = (reference_model(sd), sd)
        ref, sd = models[calibrated]
        x = fixtures.image(1, h, w)
        with torch.no_grad():
            out = ref.compress(x, rate_ind=q)
            img, z_hat, y_hat = ref.decompress(out["string_list"], beta=beta)
        eb, gc = orc.entropy_models(sd)
        o = orc.compress(sd, x, q, eb, gc)   # the oracle supplies symbols / indexes (bit-equal to the reference, see test_oracle_vs_reference)
        assert o["string_list"] == out["string_list"] and torch.equal(o["y_hat"], out["y_hat"])
        np.savez_compressed(
            os.path.join(HERE, f"codec_{name}.npz"), h=h, w=w, q=q, beta=beta, calibrated=calibrated,
            header=np.frombuffer(out["string_list"][0], dtype=np.uint8), z_string=np.frombuffer(out["string_list"][1], dtype=np.uint8),
            y_string=np.frombuffer(out["string_list"][2], dtype=np.uint8),
            y_sym=o["y_sym"].numpy().astype(np.int16), y_idx=o["y_idx"].numpy().astype(np.uint8),
            z_sym=o["z_sym"].numpy().astype(np.int16),
            y_lik=out["y_likelihood"].numpy().astype(np.float32), pred_y_bit=out["pred_y_bit"], pred_z_bit=out["pred_z_bit"],
            recon_u8=orc.to_uint8(img), psnr=orc.psnr_u8(x, img), y_hat_absmax=float(y_hat.abs().max()),
            source="reference modules from /root/reference + oracle CompressAI restatement")
        print(name, "bytes", [len(s) for s in out["string_list"]], "psnr", orc.psnr_u8(x, img))


if __name__ == "__main__":
    rans_vectors()
    if os.path.isdir(REFERENCE) and "--tables-only" not in sys.argv:
        codec_goldens()
    print("golden fixtures written to", HERE)
