"""CPU tests of the host-side mirror: config, registries, checkpoint layout, bitstream container, C ABI exports."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT


def test_config_base_merge(tmp_path, crdr_opt):
    assert crdr_opt.model_type == "BetaCondInterpCaHyperpriorCharmModel"
    assert crdr_opt.subnet.encoder.ca_kwargs.actv == "softplus" and crdr_opt.subnet.context_model.max_support_slices == 5
    assert crdr_opt.get("missing", 7) == 7 and crdr_opt["device"] == "cuda:0"
    with pytest.raises(AttributeError):
        crdr_opt.subnet.nothing
    from crdr_b200.config import BaseConfig
    (tmp_path / "a.yaml").write_text("x: {p: 1, q: 2}\ny: 3\nz: {m: 9}\n")
    (tmp_path / "b.yaml").write_text("y: 4\n")
    (tmp_path / "c.yaml").write_text("_base_: [a.yaml]\nx: {q: 5}\nz: {_delete_: true, k: 1}\n")
    cfg, _, loaded = BaseConfig._file2dict_yaml(str(tmp_path / "c.yaml"))
    assert cfg == {"x": {"p": 1, "q": 5}, "y": 3, "z": {"k": 1}} and len(loaded) == 2
    (tmp_path / "d.yaml").write_text("_base_: [a.yaml, b.yaml]\n")
    with pytest.raises(KeyError):
        BaseConfig._file2dict_yaml(str(tmp_path / "d.yaml"))


def test_registry_names_and_duplicates():
    import crdr_b200.model  # noqa: F401
    from crdr_b200 import registry as R
    for reg, name in ((R.MODEL_REGISTRY, "BetaCondInterpCaHyperpriorCharmModel"), (R.ENCODER_REGISTRY, "ElicInterpCaEncoder"),
                      (R.DECODER_REGISTRY, "ElicInterpCaBetaCondDecoder"), (R.HYPERENCODER_REGISTRY, "Minnen20HyperEncoder"),
                      (R.HYPERDECODER_REGISTRY, "Minnen20HyperDecoder"), (R.CONTEXTMODEL_REGISTRY, "Minnen20CharmContextModel"),
                      (R.ENTROPYMODEL_REGISTRY, "SteEntropyBottleneck"), (R.ENTROPYMODEL_REGISTRY, "SteGaussianMeanScaleConditional")):
        assert name in reg
    with pytest.raises(KeyError):
        R.MODEL_REGISTRY.get("Nope")
    with pytest.raises(AssertionError):
        R.MODEL_REGISTRY.register()(R.MODEL_REGISTRY.get("BetaCondInterpCaHyperpriorCharmModel"))


@pytest.fixture(scope="module")
def model(crdr_opt):
    from crdr_b200.model import build_comp_model
    torch.manual_seed(0)
    return build_comp_model(crdr_opt)


def test_state_dict_layout_matches_reference(model):
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_layout_crdr.json")))
    got = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in model.state_dict().items()}
    assert list(got) == list(want) and got == want and len(got) == 595
    assert sum(p.numel() for p in model.parameters()) == 127712963
    main, aux = model.separate_aux_parameters()
    assert list(aux) == ["entropy_model_z.quantiles"] and len(main) == 582


def test_checkpoint_roundtrip_and_setup(model, tmp_path):
    model.codec_setup()
    sd = model.state_dict()
    assert sd["entropy_model_y._quantized_cdf"].shape[0] == 64 and sd["entropy_model_z._quantized_cdf"].shape[0] == 192
    path = tmp_path / "ckpt.pth.tar"
    torch.save({"iter": 0, "comp_model": {"module." + k if i % 2 else k: v for i, (k, v) in enumerate(sd.items())}}, path)
    from crdr_b200.model import build_comp_model
    other = build_comp_model(model.opt)
    other.load_learned_weight(str(path))  # fresh module has empty tables: buffers are resized on load
    for k, v in other.state_dict().items():
        assert torch.equal(v, sd[k]), k
    assert float(model.aux_loss()) > 0


def test_no_cpu_fallback(model):
    from crdr_b200.native import NativeError
    model.device = "cpu"
    model.invalidate_engine()
    try:
        if not hasattr(model, "header_handler"):
            model.codec_setup()
            model.invalidate_engine()
        with pytest.raises(NativeError):
            model.compress(torch.zeros(1, 3, 64, 64), 0.0)
        with pytest.raises(NativeError):   # training-mode forward runs on the same engines: no CPU path either
            model.forward(torch.zeros(1, 3, 64, 64), 0.0, 0.0, is_train=True)
    finally:
        model.device = "cuda:0"


def test_header_and_container_bytes(tmp_path):
    from crdr_b200.codec_utils import MultiRateHeaderHandler, load_byte_strings, save_byte_strings
    hh = MultiRateHeaderHandler()
    hdr = hh.encode((512, 768), rate_ind=1.75, max_abs=21.9)
    assert hdr == bytes([0, 2, 0, 3, 21, 28]) and hh.decode(hdr) == {"img_size": (512, 768), "max_sample": 21, "rate_ind": 1.75}
    assert hh.encode((1, 1), y_hat=torch.tensor([-3.7, 2.0]), rate_ind=torch.tensor([4.0])) == bytes([1, 0, 1, 0, 3, 64])
    # the informational max-sample byte wraps like the reference's np.uint8 cast (numpy 1.x) instead of failing the encode
    assert hh.encode((10, 10), rate_ind=0.0, max_abs=300.0) == bytes([10, 0, 10, 0, 300 & 0xFF, 0])
    with pytest.raises(OverflowError):
        hh.encode((70000, 10), rate_ind=0.0, max_abs=1.0)
    strings = [hdr, b"", b"\x01\x02\x03" * 100]
    p = tmp_path / "x.bin"
    save_byte_strings(str(p), strings)
    raw = p.read_bytes()
    assert raw[:4] == (6).to_bytes(4, "little") and len(raw) == 4 * 3 + 6 + 300
    assert load_byte_strings(str(p)) == strings


def test_interp_gain_vectors_match_reference_formula(model):
    from crdr_b200.codec import InterpGain
    w = torch.randn(5, 1, 16, 1, 1)
    b = torch.randn(5, 1, 16, 1, 1)
    g = InterpGain(w, b, "cpu")
    for q, (l, r, a) in {0.0: (0, 1, 1.0), 1.25: (1, 2, 0.75), 4.0: (4, 4, 0.0), 3.5: (3, 4, 0.5)}.items():
        sc, sh = g.vectors(q)
        assert torch.allclose(sc, torch.nn.functional.softplus(w[l] * a + w[r] * (1 - a)).reshape(-1))
        assert torch.allclose(sh, (b[l] * a + b[r] * (1 - a)).reshape(-1))
    with pytest.raises(AssertionError):
        g.vectors(4.5)


def test_conv_lowering_matches_torch_on_cpu():
    """Tap tables / packed weights of ConvOp reproduce Conv2d and ConvTranspose2d (evaluated with plain matmuls)."""
    import torch.nn.functional as F
    from crdr_b200.engine import ConvOp

    def emulate(op, x):
        n, c, h, w = x.shape
        ho, wo = op.out_hw(h, w)
        out = torch.zeros(n, op.cout, ho, wo, dtype=torch.float64)
        xh = F.pad(x.permute(0, 2, 3, 1), (0, op.cin - c)).double()
        for ph in op.phases:
            st_in, st_out = (1, op.stride) if op.transposed else (op.stride, 1)
            hb = -(-(ho - ph.out_ph) // st_out)
            wb = -(-(wo - ph.out_pw) // st_out)
            W = ph.w_hi.double() + ph.w_lo.double() / 2048
            for bh in range(hb):
                for bw in range(wb):
                    acc = torch.zeros(n, op.cout_pad, dtype=torch.float64)
                    for t, (dh, dw) in enumerate(zip(ph.dh, ph.dw)):
                        ih, iw = bh * st_in + dh, bw * st_in + dw
                        if 0 <= ih < h and 0 <= iw < w:
                            acc += xh[:, ih, iw, :] @ W[:, t * op.cin:(t + 1) * op.cin].t()
                    out[:, :, bh * st_out + ph.out_ph, bw * st_out + ph.out_pw] = acc[:, :op.cout]
        return out + op.bias.double().view(1, -1, 1, 1)

    g = torch.Generator().manual_seed(0)
    for cin, cout, k, s, tr, hh, ww in [(8, 16, 5, 2, False, 9, 11), (16, 8, 5, 2, True, 5, 6), (16, 24, 3, 1, True, 5, 6),
                                        (8, 8, 3, 1, False, 6, 7), (8, 16, 1, 1, False, 4, 4), (3, 16, 5, 2, False, 10, 12)]:
        x = torch.randn(2, cin, hh, ww, generator=g)
        w = torch.randn(cin, cout, k, k, generator=g) if tr else torch.randn(cout, cin, k, k, generator=g)
        b = torch.randn(cout, generator=g)
        op = ConvOp(w, b, transposed=tr, stride=s, padding=k // 2, output_padding=(s - 1 if tr else 0),
                    cin_pad=8 if cin == 3 else None, device="cpu")
        ref = (F.conv_transpose2d(x, w, b, stride=s, padding=k // 2, output_padding=s - 1) if tr
               else F.conv2d(x, w, b, stride=s, padding=k // 2))
        got = emulate(op, x)
        assert got.shape == ref.shape and (got - ref.double()).abs().max() < 1e-4


def _declared_symbols(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(crdr_[a-z0-9_]+)\s*\(", text)))


def test_c_abi_libraries_export_every_declared_symbol():
    """No compute calls here (no GPU): only that the shared objects load and export the header's surface."""
    from crdr_b200 import native, rans
    from crdr_b200.build import build_all
    build_all()
    for so, header, listed in ((native.SM100_SO, "crdr_b200.h", native.SM100_SYMBOLS), (native.RANS_SO, "crdr_rans.h", rans.RANS_SYMBOLS)):
        declared = _declared_symbols(header)
        assert declared == sorted(listed), (header, set(declared) ^ set(listed))
        lib = ctypes.CDLL(so)
        for name in declared:
            assert hasattr(lib, name), f"{so} does not export {name}"
    lib = ctypes.CDLL(native.SM100_SO)
    assert lib.crdr_abi_version() == 1
    # struct sizes on the Python side must agree with the C header (compiled here with gcc)
    import subprocess, tempfile
    src = '#include <stdio.h>\n#include "crdr_b200.h"\nint main(){printf("%zu %zu %zu %zu", sizeof(crdr_planes), sizeof(crdr_conv_desc), sizeof(crdr_gauss_desc), sizeof(crdr_eb_desc));return 0;}'
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(td, "s"), os.path.join(td, "s.c")])
        sizes = [int(v) for v in subprocess.check_output([os.path.join(td, "s")]).split()]
    assert sizes == [ctypes.sizeof(native.Planes), ctypes.sizeof(native.ConvDesc), ctypes.sizeof(native.GaussDesc), ctypes.sizeof(native.EbDesc)]


def test_chunk_split_and_pipeline_scheduler(model):
    """Host logic of the codec API's software pipeline (model.py `_chunks` / `_drive`): chunk boundaries cover the batch
    exactly once, and generators are resumed round-robin, each only after its own event has been synchronised."""
    cls = type(model)
    assert model._chunks(3) == [(0, 3)]                                     # below pipeline_min_images: one chunk
    # compress_batch: 3:1 (a short last chunk shortens the device span before the exposed last host encode);
    # decompress_batch: 2:1 (a short last chunk shortens the pipeline drain)
    wc = model.pipeline_weights_compress
    assert model._chunks(24, None, wc) == [(0, 18), (18, 24)] and model._chunks(9, None, wc) == [(0, 7), (7, 9)]
    assert model._chunks(24, None, (1.0, 1.0)) == [(0, 12), (12, 24)] and model._chunks(9, None, (1.0, 1.0)) == [(0, 4), (4, 9)]
    assert model._chunks(24) == [(0, 16), (16, 24)] and model._chunks(9) == [(0, 6), (6, 9)]
    # the shape the codec calls use: the defaults above with enough coder threads, the few-thread profile (coder paced:
    # compress 2:1, decompress 1:1:1) on a small slice of the host, and never when the instance sets a shape itself
    saved_few = cls.pipeline_few_threads
    try:
        cls.pipeline_few_threads = 0
        assert model._pipeline_shape(True) == (2, wc) and model._pipeline_shape(False) == (2, model.pipeline_weights)
        cls.pipeline_few_threads = 10 ** 6
        assert model._pipeline_shape(True) == (2, (2.0, 1.0)) and model._pipeline_shape(False) == (3, (1.0, 1.0, 1.0))
        assert model._chunks(24, *model._pipeline_shape(False)) == [(0, 8), (8, 16), (16, 24)]
        model.pipeline_chunks = 2
        assert model._pipeline_shape(False) == (2, model.pipeline_weights)
        del model.pipeline_chunks
    finally:
        cls.pipeline_few_threads = saved_few
    saved = model.pipeline_weights
    model.pipeline_chunks, model.pipeline_weights = 3, (3, 2, 1)
    try:
        ch = model._chunks(24)
        assert ch == [(0, 12), (12, 20), (20, 24)]
        model.pipeline_weights = None
        for n in (8, 10, 25):
            ch = model._chunks(n)
            assert ch[0][0] == 0 and ch[-1][1] == n and all(a[1] == b[0] for a, b in zip(ch, ch[1:])) and len(ch) == 3
    finally:
        model.pipeline_chunks, model.pipeline_weights = 2, saved

    log = []

    class Ev:
        def __init__(self, name):
            self.name = name

        def synchronize(self):
            log.append(("sync", self.name))

    def gen(name, stages):
        for s in range(stages):
            log.append(("enqueue", name, s))
            yield Ev(f"{name}{s}")
            log.append(("host", name, s))
        return name.upper()

    out = cls._drive([gen("a", 2), gen("b", 3), gen("c", 0)])
    assert out == ["A", "B", "C"]
    # both first device segments are enqueued before any host stage runs; afterwards strict round-robin, and a host
    # stage never runs before its event was synchronised
    assert log[:2] == [("enqueue", "a", 0), ("enqueue", "b", 0)]
    for i, rec in enumerate(log):
        if rec[0] == "host":
            assert log[i - 1] == ("sync", f"{rec[1]}{rec[2]}")
    hosts = [r[1:] for r in log if r[0] == "host"]
    assert hosts == [("a", 0), ("b", 0), ("a", 1), ("b", 1), ("b", 2)]


def test_trainer_plugins_and_loss_config():
    """The reference's trainer / discriminator registry names resolve (src/trainer/__init__.py:9-26,
    src/models/discriminator/__init__.py:15-30) and the loss / optimiser sections of the stage-2 / stage-3 yaml files map
    onto the lowered step's settings."""
    import src  # noqa: F401
    from crdr_b200.config import BaseConfig
    from crdr_b200.registry import DISCRIMINATOR_REGISTRY, TRAINER_REGISTRY
    from crdr_b200.trainers import _loss_kwargs
    from src.models.discriminator import build_discriminator
    from src.trainer import build_trainer  # noqa: F401
    assert {"RateDistortionTrainer", "MultirateBetaCondHrrGanRateDistortionTrainer"} <= set(TRAINER_REGISTRY.keys())
    assert {"CLIC21GVAEDiscriminator", "ModuleListDiscriminator"} <= set(DISCRIMINATOR_REGISTRY.keys())
    o2 = BaseConfig.fromfile(os.path.join(ROOT, "config", "crdr_stage_2.yaml"), device="cpu", is_train=True)
    k2 = _loss_kwargs(o2)
    assert k2["rate_lambda_a"] == [3.6, 1.8, 0.8, 0.4, 0.1] and k2["rate_lambda_b"] == 2.0 ** -6 and k2["lambda_mse"] == 150.0
    assert k2["clip_max_norm"] == 1.0 and k2["lr"] == 1e-4 and k2["aux_lr"] == 1e-3 and k2["perceptual_weight"] == 1.0
    o3 = BaseConfig.fromfile(os.path.join(ROOT, "config", "crdr_stage_3.yaml"), device="cpu", is_train=True)
    assert o3["trainer"]["type"] == "MultirateBetaCondHrrGanRateDistortionTrainer" and _loss_kwargs(o3)["target_rate"] == [0.0] * 5
    d = build_discriminator(dict(o3["discriminator"]))
    assert len(d.subD_list) == 5 and sum(p.numel() for p in d.subD_list[0].parameters()) == 4689985


def test_weight_index_maps_reproduce_host_packing():
    """Training re-packs every tensor-core matrix on the device by a gather through an index map (crdr_pack_weights).  The
    maps come from pushing an index tensor through the host packing code; here the gather is emulated with torch indexing
    and must reproduce the host-packed matrices for every kind of convolution of the path: plain / strided / transposed,
    two input ranges, the im2col'd first layer, the phase-packed last layer, and their dgrad adjoints (value check against
    an explicit transposed / flipped weight)."""
    import torch
    from crdr_b200 import backward as bw
    from crdr_b200.codec import AnalysisEngine, SynthesisEngine
    from crdr_b200.engine import ConvOp, LO_SCALE

    def emulate(pc):
        flat = pc.master.reshape(-1)
        for phs in pc.op.phases:
            idx = phs.map.long()
            v = torch.where(idx >= 0, flat[idx.clamp_min(0)], torch.zeros(()))
            hi = v.half()
            yield hi, ((v - hi.float()) * LO_SCALE).half()

    g = torch.Generator().manual_seed(0)
    cases = [
        (torch.randn(96, 192, 1, 1, generator=g), None, {}),
        (torch.randn(192, 192, 5, 5, generator=g) * 0.1, None, dict(stride=2, padding=2)),
        (torch.randn(320, 256, 5, 5, generator=g) * 0.1, None, dict(transposed=True, stride=2, padding=2, output_padding=1)),
        (torch.randn(224, 352, 5, 5, generator=g) * 0.1, None, dict(padding=2, seg_lens=[320, 32])),
        (torch.randn(192, 3, 5, 5, generator=g), AnalysisEngine._patch_weight, {}),
        (torch.randn(256, 3, 5, 5, generator=g), SynthesisEngine._phase_weight, dict(padding=1)),
    ]
    for w, tf, kw in cases:
        host = ConvOp(tf(w) if tf else w, None, device="cpu", **kw)
        pc = bw.PackedConv(w, tf or (lambda t: t), two_planes=True, **kw)
        for phs, (hi, lo) in zip(host.phases, emulate(pc)):
            assert torch.equal(phs.w_hi, hi) and torch.equal(phs.w_lo, lo)
    # adjoints: the dgrad matrices of a 3x3 stride-1 convolution hold the transposed, flipped weights
    w = torch.randn(96, 64, 3, 3, generator=g)
    ds = bw.DgradSet(w, False, 1, 1, 3, [(0, 64)])
    host = ConvOp(w.permute(1, 0, 2, 3).flip(2, 3).contiguous(), None, device="cpu", padding=1)
    (pc, coff, cnt), = ds.parts
    assert (coff, cnt) == (0, 64)
    for phs, (hi, _) in zip(host.phases, emulate(pc)):
        assert torch.equal(phs.w_hi, hi)
    # more than 320 input channels: two chunks, written to consecutive channel ranges of the input gradient
    ds = bw.DgradSet(torch.randn(224, 480, 5, 5, generator=g), False, 1, 2, 5, [(320, 320), (960, 160)])
    assert [(c, n) for _, c, n in ds.parts] == [(320, 320), (960, 160)]
