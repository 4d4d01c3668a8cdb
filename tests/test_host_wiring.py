"""Host logic above the C ABI, on a machine WITHOUT a GPU: the drop-in modules (Wav2Letter, Jasper, CTCLoss) run their real
autograd Functions, descriptor geometry, halo / mask / unfold / residual wiring and length bookkeeping, with every C-ABI call
answered by the torch restatement in tests/_host_sim.py (test infrastructure; the product itself has no CPU path).  Results are
held to the same fixtures frozen from the unmodified reference, and the same tolerances, as the `-m gpu` model tests: a wiring
error (a dropped residual gradient, a wrong row offset, a halo folded onto the wrong row) shows up as an error of order 1."""
import json

import numpy as np
import pytest
import torch

import _host_sim
import _kernel_emu
import _layerwise as L
from oracle import w2l_oracle as O
from test_gpu_models import E2E_TOY_BOUND, TOL_EMU, TOL_REF

# "sim": every C-ABI call answered by the torch restatement of tests/_host_sim.py; "emu": every call runs the library's own
# kernel source on the host (tests/_emu_backend.py) -- the tcgen05 implicit GEMMs included, through their own C wrappers
# "cabi": the unmodified functional.py over the library's own extern "C" wrappers + kernels compiled for the host (tests/_emu_cabi.py)
_NEEDS_GXX = pytest.mark.skipif(not _kernel_emu.available(), reason="needs g++ and the CUDA headers")
BACKENDS = ["sim", pytest.param("emu", marks=_NEEDS_GXX), pytest.param("cabi", marks=_NEEDS_GXX)]


def _gemm_launches(backend):
    if backend == "cabi":
        import _emu_cabi
        return int(_emu_cabi.builds()[-1].lib.emu_launch_count())
    if backend != "emu":
        return 0
    import _emu_backend
    return int(_emu_backend.gemm().lib.emu_launch_count())


def _install(monkeypatch, backend):
    if backend == "cabi":
        import _emu_cabi
        return _emu_cabi.install(monkeypatch)
    if backend == "emu":
        import _emu_backend
        return _emu_backend.install(monkeypatch)
    return _host_sim.install(monkeypatch)


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().flatten(), torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _load_sd(model, g, prefix):
    sd = {k[len(prefix):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix)}
    missing, unexpected = model.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return sd


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("fixture", ["w2l_small", "w2l_strided", "w2l_narrow", "w2l_odd"])
def test_w2l_wiring_against_reference_fixture(golden, monkeypatch, fixture, backend):
    from wav2letter_pytorch_b200 import config
    from wav2letter_pytorch_b200.wav2letter import Wav2Letter
    _install(monkeypatch, backend)
    launches0 = _gemm_launches(backend)
    from wav2letter_pytorch_b200.layers import FusedBnReduce
    fused0 = FusedBnReduce.fused_launches
    g = golden(fixture)
    layers = [dict(output_size=int(o), kernel_size=int(k), stride=int(s), dilation=int(d), dropout=-1) for o, k, s, d in g["layers"]]
    cfg = config.compose(overrides=["model.mid_layers=3"]).model
    cfg["layers"] = config.to_attr(layers)
    if fixture == "w2l_odd":                      # 161 STFT bins (input_size unset), hidden widths 250 / 36 / 250: padded internally
        cfg["input_size"] = 0
    model = Wav2Letter(cfg)
    assert sorted(model.state_dict().keys()) == sorted(k[4:] for k in g.files if k.startswith("sd0:"))
    _load_sd(model, g, "sd0:")
    model.train()
    x, il, tg, tl = (torch.from_numpy(g[k]) for k in ("x", "il", "tg", "tl"))
    hs, out, ol, loss = L.w2l_run_blocks(model, x, il, tg, tl)
    assert out.shape == tuple(g["train:out"].shape) and out.dtype == torch.float32 and out.is_contiguous()
    assert ol.dtype == il.dtype and np.array_equal(ol.numpy(), g["train:out_len"])
    assert rel_l2(out.detach(), g["train:out"]) < 2e-2
    assert abs(loss.item() - float(g["train:loss"])) < 2e-2 * abs(float(g["train:loss"]))
    # the GPU tests' closed tolerances (tests/test_gpu_models.py): every block on its own, then one fixed end-to-end bound
    table = L.w2l_layerwise_table(model, hs, out)
    bad = L.check_table(table, TOL_EMU, TOL_REF)
    assert not bad, (bad, L.format_table(table))
    for name, p in model.named_parameters():
        ref = g["train:grad:" + name]
        assert p.grad is not None and p.grad.shape == p.shape, name
        if name.endswith("conv1.bias") and "conv1d_3" not in name:      # analytically zero under train-mode BN
            assert p.grad.abs().max().item() == 0.0
            continue
        assert rel_l2(p.grad, ref) < E2E_TOY_BOUND, (name, rel_l2(p.grad, ref))
    sd1 = {k[4:]: g[k] for k in g.files if k.startswith("sd1:")}
    for k, v in model.state_dict().items():
        if "running" in k:
            np.testing.assert_allclose(v.numpy(), sd1[k], rtol=2e-2, atol=2e-3, err_msg=k)
        if "num_batches_tracked" in k:
            assert int(v) == int(sd1[k])
    _load_sd(model, g, "sd1:")
    model.eval()
    with torch.no_grad():
        o, ol = model(x, il)
    assert rel_l2(o, g["eval:out"]) < 2e-2 and np.array_equal(ol.numpy(), g["eval:out_len"])
    assert model.scaling_factor == int(g["scaling_factor"])
    if backend != "sim":                          # the GEMMs really went through the emulated conv_gemm_kernel: fwd + dgrad + wgrad per layer
        assert _gemm_launches(backend) - launches0 >= 3 * len(layers)
    # every BatchNorm block whose output feeds the next layer unchanged had its backward reduction folded into that layer's
    # backward-data GEMM (FusedBnReduce; a strided inner layer unfolds its input first: that producer keeps the separate pass)
    want = 0 if backend == "sim" else sum(1 for i in range(len(layers)) if i + 1 >= len(layers) or layers[i + 1]["stride"] == 1)
    assert FusedBnReduce.fused_launches - fused0 == want, (FusedBnReduce.fused_launches - fused0, want)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("fixture,seed", [("jasper_dense", 4), ("jasper_small", 2), ("jasper_strided", 10), ("jasper_odd", 23)])
def test_jasper_wiring_against_reference_fixture(golden, monkeypatch, fixture, seed, backend):
    from wav2letter_pytorch_b200 import config
    from wav2letter_pytorch_b200.jasper import Jasper
    _install(monkeypatch, backend)
    from wav2letter_pytorch_b200.layers import FusedBnReduce
    fused0 = FusedBnReduce.fused_launches
    g = golden(fixture)
    blocks = [dict(b, dropout=0) for b in json.loads(str(g["blocks_json"]))]
    cfg = config.compose(overrides=["model=jasper", "model.mid_layers=%d" % len(blocks)]).model
    cfg["jasper_blocks"] = config.to_attr(blocks)
    if fixture == "jasper_odd":                   # 161 STFT bins (input_size unset), widths 100 / 36 / 250 / 52: padded internally
        cfg["input_size"] = 0
    torch.manual_seed(seed)
    model = Jasper(cfg)
    for k in g.files:                                                   # seeded construction == the reference's
        if k.startswith("sd_init:"):
            assert np.array_equal(model.state_dict()[k[8:]].numpy(), g[k]), k
    assert sorted(model.state_dict().keys()) == sorted(k[4:] for k in g.files if k.startswith("sd0:"))
    _load_sd(model, g, "sd0:")
    model.train()
    x, il, tg, tl = (torch.from_numpy(g[k]) for k in ("x", "il", "tg", "tl"))
    hs, taps, rows, out, ol, loss = L.jasper_run_blocks(model, x, il, tg, tl)
    assert np.array_equal(ol.numpy(), g["train:out_len"]) and ol.dtype == torch.int64
    assert rel_l2(out.detach(), g["train:out"]) < 2e-2
    assert abs(loss.item() - float(g["train:loss"])) < 2e-2 * abs(float(g["train:loss"]))
    # the GPU tests' closed tolerances (tests/test_gpu_models.py): every conv+BN group on its own, then one fixed end-to-end bound
    table = L.jasper_layerwise_table(model, O.jasper_block_specs(blocks), hs, taps, rows, out)
    bad = L.check_table(table, TOL_EMU, TOL_REF)
    assert not bad, (bad, L.format_table(table))
    for name, p in model.named_parameters():
        ref = torch.from_numpy(g["train:grad:" + name])
        assert p.grad is not None and p.grad.shape == p.shape, name
        assert rel_l2(p.grad, ref) < E2E_TOY_BOUND, (name, rel_l2(p.grad, ref))
    if backend == "sim":
        assert FusedBnReduce.fused_launches == fused0
    elif fixture == "jasper_dense":               # repeats inside a block chain linearly: folded; a block output feeds two consumers: not
        assert FusedBnReduce.fused_launches - fused0 > 0
    for k in g.files:
        if k.startswith("sd1:") and "running" in k:
            np.testing.assert_allclose(model.state_dict()[k[4:]].numpy(), g[k], rtol=2e-2, atol=2e-3, err_msg=k)
    model.eval()
    with torch.no_grad():
        o, ol = model(x, il)
    assert rel_l2(o, g["eval:out"]) < 2e-2 and abs(float(o.sum(-1).mean()) - 1.0) < 1e-5    # probabilities in eval


def test_strided_jasper_block_with_residual_raises(monkeypatch):
    """jasper.py:412: `out + res_out` cannot be formed when the block strides (the 1x1 residual conv does not)"""
    from wav2letter_pytorch_b200 import config
    from wav2letter_pytorch_b200.jasper import Jasper
    _host_sim.install(monkeypatch)
    blocks = [dict(layer_size=64, kernel_size=5, stride=1, residual=False, separable=False, repeat=1, dropout=0),
              dict(layer_size=64, kernel_size=5, stride=2, residual=True, separable=False, repeat=1, dropout=0)]
    cfg = config.compose(overrides=["model=jasper", "model.mid_layers=2"]).model
    cfg["jasper_blocks"] = config.to_attr(blocks)
    model = Jasper(cfg).train()
    with pytest.raises(RuntimeError):
        model(torch.randn(2, 64, 50), torch.tensor([50, 40]))


def test_head_block_standalone(monkeypatch):
    """Conv1dBlock(..., bn=False, activation_use=False) called on its own (wav2letter.py:40-47, 69): conv + bias, NCW in / NCW out"""
    from wav2letter_pytorch_b200.wav2letter import Conv1dBlock
    _host_sim.install(monkeypatch)
    torch.manual_seed(0)
    blk = Conv1dBlock(64, 29, (1,), 1, bn=False, activation_use=False)
    x = torch.randn(2, 64, 37)
    y = blk(x)
    w, b = blk.conv1.weight.detach(), blk.conv1.bias.detach()
    want = torch.nn.functional.conv1d(x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), b)
    assert y.shape == (2, 29, 37) and y.dtype == torch.float32
    assert rel_l2(y.detach(), want) < 1e-5
    g = torch.randn(y.shape)
    y.backward(g)
    wref = w.clone().requires_grad_(True)
    bref = b.clone().requires_grad_(True)
    torch.nn.functional.conv1d(x.to(torch.bfloat16).float(), wref, bref).backward(g)
    assert rel_l2(blk.conv1.weight.grad, wref.grad) < 1e-2 and rel_l2(blk.conv1.bias.grad, bref.grad) < 1e-2


@pytest.mark.skipif(not _kernel_emu.available(), reason="needs g++ and the CUDA headers")
def test_training_step_and_decode_over_emulated_abi(monkeypatch):
    """what __graft_entry__.smoke() does on cuda:0, on the host over the emulated C ABI: training_step (conv stack, CTC loss from the
    library's kernels, greedy decode + WER/CER for the log), backward, then an eval forward whose transcripts must equal the oracle's
    greedy decode of the same scores bit for bit"""
    from wav2letter_pytorch_b200 import config
    from wav2letter_pytorch_b200.wav2letter import Wav2Letter
    _install(monkeypatch, "cabi")
    torch.manual_seed(0)
    cfg = config.compose(overrides=["model.mid_layers=3", "optimizer=novograd"]).model
    for l in cfg.layers:
        l["dropout"] = 0.0
    model = Wav2Letter(cfg).train()
    model._optimizer = torch.optim.SGD(model.parameters(), lr=0.0)              # training_step logs the learning rate
    x, il, tg, tl = O.synthetic_batch(3, 1, seed=1, ragged=True)
    texts = ["".join(O.ENGLISH_LOWERCASE[c] for c in row[: int(n)].tolist()) for row, n in zip(tg, tl)]
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    loss = model.training_step((x, il, tg, tl, None, texts), 0)
    loss.backward()
    assert set(model.logged) >= {"train_loss", "train_cer", "train_wer", "train_len_ratio", "learning_rate"}
    specs = O.w2l_layer_specs(3, dropout=0.0)
    lp, ol = O.w2l_forward_bf16emu(x, il, sd0, specs, training=True)
    ref = torch.nn.CTCLoss(blank=0, reduction="mean", zero_infinity=True)(lp.transpose(0, 1), tg, ol, tl)
    assert abs(loss.item() - ref.item()) < 5e-3 * abs(ref.item()), (loss.item(), ref.item())
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
    model.eval()
    with torch.no_grad():
        out, out_len = model(x, il)
    hyp = model.ctc_decoder.decode(out, out_len)
    want, _ = O.greedy_decode(out.numpy(), out_len.numpy())
    assert hyp == want


def test_layerwise_parity_tables_sim(monkeypatch):
    """tests/_layerwise.py (the teacher-forced per-block parity of tests/test_gpu_parity_headline.py) on the host backend at a small
    size: the tables must come out under the GPU test's fixed bounds here too -- which pins the helper itself (row bookkeeping of the
    taps, residual + first-group input gradients, masks) before it meets the hardware"""
    import _layerwise as L
    from wav2letter_pytorch_b200 import config, functional as F
    from wav2letter_pytorch_b200.jasper import Jasper
    from wav2letter_pytorch_b200.wav2letter import Wav2Letter
    _install(monkeypatch, "sim")
    cfg = config.compose(overrides=["model.mid_layers=3"]).model
    for l in cfg.layers:
        l["dropout"] = 0.0
    torch.manual_seed(0)
    model = Wav2Letter(cfg).train()
    x, il, tg, tl = O.synthetic_batch(2, 1, seed=5, ragged=True)
    hs, out, ol, loss = L.w2l_run_blocks(model, x, il, tg, tl)
    table = L.w2l_layerwise_table(model, hs, out)
    assert len(table) == 4 and not L.check_table(table, TOL_EMU, TOL_REF), L.format_table(table)
    blocks = [dict(layer_size=64, kernel_size=11, stride=2, dilation=1, residual=False, repeat=1, separable=False, dropout=0.0),
              dict(layer_size=128, kernel_size=5, stride=1, dilation=1, residual=True, repeat=3, separable=False, dropout=0.0),
              dict(layer_size=64, kernel_size=3, stride=1, dilation=2, residual=False, repeat=1, separable=False, dropout=0.0)]
    cfg = config.compose(overrides=["model=jasper", "model.mid_layers=3"]).model
    cfg["jasper_blocks"] = config.to_attr(blocks)
    torch.manual_seed(0)
    model = Jasper(cfg).train()
    x, il, tg, tl = O.synthetic_batch(2, 2, seed=6, ragged=True)
    hs, taps, rows, out, ol, loss = L.jasper_run_blocks(model, x, il, tg, tl, F)
    table = L.jasper_layerwise_table(model, O.jasper_block_specs(blocks), hs, taps, rows, out)
    assert len(table) == 6 and not L.check_table(table, TOL_EMU, TOL_REF), L.format_table(table)
    # a wiring error must show: drop the residual branch's contribution from the block's input gradient
    hs[1].grad.mul_(0.9)
    table = L.jasper_layerwise_table(model, O.jasper_block_specs(blocks), hs, taps, rows, out)
    assert any(q == "d_input" for _, q, _, _ in L.check_table(table, TOL_EMU, TOL_REF))
