"""Model-level parity on a B200: the drop-in modules (Wav2Letter, Conv1dBlock, CTCLoss, GreedyDecoder, Novograd)
against the reference's frozen outputs (tests/golden, produced by the unmodified reference) and the CPU oracle.

Tolerances (stated per SURVEY 8c): the conv path computes with bf16 operands / fp32 accumulation and stores bf16
activations, so logits are compared by relative L2 <= 2e-2 and gradients by relative L2 <= 6e-2 against the fp32
reference; the CTC loss kernel itself is fp32 (<= 1e-4 relative on identical log-probs); transcripts are bit-exact
on identical scores."""
import json

import numpy as np
import pytest
import torch

import _layerwise as L
from oracle import w2l_oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

# fixed bounds: teacher-forced per-block parity vs the bf16-emulating oracle / vs the fp32 oracle, and the one
# end-to-end bound for the toy fixtures' gradients against the reference's frozen ones
TOL_EMU = {"out": 5e-3, "d_input": 2e-2, "d_param": 2e-2}
TOL_REF = {"out": 1e-2, "d_input": 1.5e-1, "d_param": 1.5e-1}
E2E_TOY_BOUND = 0.35


@pytest.fixture(scope="module")
def pkg():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import wav2letter_pytorch_b200 as p
    return p


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().flatten().cpu(), torch.as_tensor(b).double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _cfg(pkg, layers, mid):
    from wav2letter_pytorch_b200 import config
    cfg = config.compose(overrides=["model.mid_layers=%d" % mid])
    cfg.model["layers"] = config.to_attr(layers)
    return cfg.model


def _load_sd(model, g, prefix):
    sd = {k[len(prefix):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix)}
    missing, unexpected = model.load_state_dict(sd, strict=True)
    assert not missing and not unexpected


def test_w2l_golden_train_eval(pkg, golden):
    check_w2l_golden(pkg, golden("w2l_small"))


def test_w2l_golden_with_fused_bn_reduce(pkg, golden, monkeypatch):
    """the same fixture with the opt-in W2L_FUSE_BN_REDUCE path: every block's backward reduction comes out of the next layer's
    backward-data GEMM epilogue (layers.FusedBnReduce) -- same closed tolerances, and the number of folded reductions is the number of
    BatchNorm blocks"""
    from wav2letter_pytorch_b200.layers import FusedBnReduce
    monkeypatch.setattr(FusedBnReduce, "enabled", True)
    before = FusedBnReduce.fused_launches
    g = golden("w2l_small")
    check_w2l_golden(pkg, g)
    assert FusedBnReduce.fused_launches - before == len(g["layers"])


def test_w2l_odd_widths_golden(pkg, golden):
    """channel counts the reference accepts and the tensor-core path pads internally: 161 STFT bins in (input_size unset,
    wav2letter.py:52-56), hidden widths 250 / 36 / 250 (wav2letter.py:59-64; 250 is the original paper's) -- same checks and bounds as
    the other fixtures; the per-block table also asserts that the surplus channels of every buffer and gradient are exact zeros"""
    check_w2l_golden(pkg, golden("w2l_odd"), input_size=0)


def check_w2l_golden(pkg, g, input_size=None):
    """train step + eval forward of a small Wav2Letter against a fixture frozen from the unmodified reference"""
    from wav2letter_pytorch_b200.wav2letter import Wav2Letter
    layers = [dict(output_size=int(o), kernel_size=int(k), stride=int(s), dilation=int(d), dropout=-1) for o, k, s, d in g["layers"]]
    cfg = _cfg(pkg, layers, len(layers))
    if input_size is not None:
        cfg["input_size"] = input_size
    model = Wav2Letter(cfg)
    head = "conv1d_%d" % len(layers)                        # the bias-only label head follows the BatchNorm blocks
    assert sorted(model.state_dict().keys()) == sorted(k[4:] for k in g.files if k.startswith("sd0:"))     # checkpoint contract
    _load_sd(model, g, "sd0:")
    model.cuda().train()
    x, il = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["il"]).cuda()
    tg, tl = torch.from_numpy(g["tg"]).cuda(), torch.from_numpy(g["tl"]).cuda()
    # ONE training forward (the running statistics are compared below) through the model's own forward(), with its parity tap on
    hs, out, ol, loss = L.w2l_run_blocks(model, x, il, tg, tl)
    assert out.shape == tuple(g["train:out"].shape) and out.dtype == torch.float32 and out.is_contiguous()
    assert ol.dtype == il.dtype and np.array_equal(ol.cpu().numpy(), g["train:out_len"])
    assert rel_l2(out, g["train:out"]) < 2e-2
    assert abs(loss.item() - float(g["train:loss"])) < 2e-2 * abs(float(g["train:loss"]))
    # ---- gradients, closed tolerances: every block on its own (teacher-forced: the oracle block gets the activations / gradients the
    # device fed into that block, tests/_layerwise.py) against the oracle with the device's bf16 storage emulated AND against plain fp32
    table = L.w2l_layerwise_table(model, hs, out)
    bad = L.check_table(table, TOL_EMU, TOL_REF)
    assert not bad, (bad, L.format_table(table))
    # ---- end to end against the gradients the unmodified reference produced (frozen in the fixture): one fixed bound.  It is loose
    # because a freshly initialised BatchNorm stack amplifies every rounding by ~1.2x per layer (the host emulation of the same
    # bf16 storage sits 0.1-0.2 from fp32 on these toy models); the per-block table above is what catches a wiring error
    for name, p in model.named_parameters():
        ref = g["train:grad:" + name]
        assert p.grad is not None and p.grad.shape == p.shape, name
        if name.endswith("conv1.bias") and head not in name:           # analytically zero under train-mode BN
            assert p.grad.abs().max().item() == 0.0
            continue
        err_ref = rel_l2(p.grad, ref)
        assert err_ref < E2E_TOY_BOUND, (name, err_ref)
    sd1 = {k[4:]: g[k] for k in g.files if k.startswith("sd1:")}
    for k, v in model.state_dict().items():
        if "running" in k:
            np.testing.assert_allclose(v.cpu().numpy(), sd1[k], rtol=2e-2, atol=2e-3, err_msg=k)
        if "num_batches_tracked" in k:
            assert int(v) == int(sd1[k])
    # ---- eval mode (BatchNorm folded into the conv epilogue)
    _load_sd(model, g, "sd1:")
    model.eval()
    with torch.no_grad():
        o, ol = model(x, il)
    assert rel_l2(o, g["eval:out"]) < 2e-2
    # transcripts: bit-exact on identical scores; end-to-end agreement reported separately
    dec = model.ctc_decoder.decode(torch.from_numpy(g["eval:out"]).cuda(), torch.from_numpy(g["eval:out_len"]).cuda())
    assert dec == [str(s) for s in g["eval:decoded"]]
    assert model.scaling_factor == int(g["scaling_factor"])


# fp32-faithful mode (precision="tf32"): activations / weights / gradients fp32 in memory, tf32 multiplies, fp32 accumulation.
# Bounds against the UNMODIFIED reference's frozen fp32 outputs (host emulation of the same kernels: out 1.2e-4, loss 1.4e-5, worst
# parameter gradient 5.5e-2 -- the first layer's weights of this freshly initialised BatchNorm stack, where the bf16 path is
# allowed 0.35 -- running statistics 1.1e-3): logits 20x, loss 200x, gradients 3.5x inside the bf16 path's bounds
TOL_TF32 = {"out": 1e-3, "loss": 1e-4, "grad": 1e-1, "running": 5e-3}


@pytest.mark.parametrize("fixture", ["w2l_small", "w2l_odd"])
def test_w2l_golden_fp32_faithful_mode(pkg, golden, fixture):
    """SURVEY 8c asks for logits / gradients vs torch fp32: the same fixture as above with the model in precision='tf32' (fp32 storage,
    kind::tf32 GEMMs incl. the weight gradient over transposed operands, fp32 BatchNorm / activation passes), train step + eval;
    ``w2l_odd``: the same mode over internally padded channel counts (161 -> 250 -> 36 -> 250)"""
    from wav2letter_pytorch_b200.wav2letter import Wav2Letter
    g = golden(fixture)
    layers = [dict(output_size=int(o), kernel_size=int(k), stride=int(s), dilation=int(d), dropout=-1) for o, k, s, d in g["layers"]]
    cfg = _cfg(pkg, layers, len(layers))
    cfg["precision"] = "tf32"
    if fixture == "w2l_odd":
        cfg["input_size"] = 0
    model = Wav2Letter(cfg)
    assert model.precision == "tf32" and all(m.conv1.f32 for m in model.conv1ds.children())
    head = "conv1d_%d" % len(layers)
    _load_sd(model, g, "sd0:")
    model.cuda().train()
    x, il = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["il"]).cuda()
    tg, tl = torch.from_numpy(g["tg"]).cuda(), torch.from_numpy(g["tl"]).cuda()
    out, ol = model(x, il)
    loss = model.criterion(out.transpose(0, 1), tg, ol, tl)
    loss.backward()
    report = {"out": rel_l2(out, g["train:out"]), "loss": abs(loss.item() - float(g["train:loss"])) / abs(float(g["train:loss"]))}
    assert out.dtype == torch.float32 and np.array_equal(ol.cpu().numpy(), g["train:out_len"])
    worst = ("", 0.0)
    for name, p in model.named_parameters():
        ref = g["train:grad:" + name]
        assert p.grad is not None and p.grad.shape == p.shape and p.grad.dtype == torch.float32, name
        if name.endswith("conv1.bias") and head not in name:
            assert p.grad.abs().max().item() == 0.0
            continue
        e = rel_l2(p.grad, ref)
        if e > worst[1]:
            worst = (name, e)
    report["grad"] = worst[1]
    sd1 = {k[4:]: g[k] for k in g.files if k.startswith("sd1:")}
    run = 0.0
    for k, v in model.state_dict().items():
        if "running" in k:
            run = max(run, rel_l2(v, sd1[k]))
    report["running"] = run
    print("tf32 mode vs the reference fixture:", json.dumps(report), "worst gradient:", worst[0])
    for key, bound in TOL_TF32.items():
        assert report[key] < bound, (key, report, worst)
    _load_sd(model, g, "sd1:")
    model.eval()
    with torch.no_grad():
        o, _ = model(x, il)
    assert rel_l2(o, g["eval:out"]) < TOL_TF32["out"]
    # end-to-end transcripts of this random-init model (argmax near-ties everywhere): >= 95 % of the characters agree with the
    # reference's (bit-exactness holds on identical scores, test_decode_golden)
    import difflib
    dec = model.ctc_decoder.decode(o, torch.from_numpy(g["eval:out_len"]).cuda())
    for a, b in zip(dec, [str(s) for s in g["eval:decoded"]]):
        assert difflib.SequenceMatcher(None, a, b).ratio() >= 0.95, (a, b)


def test_conv1d_block_golden(pkg, golden):
    """Conv1dBlock(64, 256, (11,), 2): the reference's first layer, through the standalone NCW interface."""
    from wav2letter_pytorch_b200.wav2letter import Conv1dBlock
    g = golden("conv_block")
    blk = Conv1dBlock(64, 256, (11,), 2, drop_out_prob=0.0)
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("l0:") and k[3:] in blk.state_dict()}
    blk.load_state_dict(sd)
    blk.cuda().train()
    x = torch.from_numpy(g["l0:x"]).cuda()
    y = blk(x)
    assert y.shape == tuple(g["l0:y_train"].shape)
    assert rel_l2(y, g["l0:y_train"]) < 1.5e-2
    np.testing.assert_allclose(blk.batch_norm.running_mean.cpu().numpy(), g["l0:running_mean_after"], rtol=2e-2, atol=2e-3)
    np.testing.assert_allclose(blk.batch_norm.running_var.cpu().numpy(), g["l0:running_var_after"], rtol=2e-2, atol=2e-3)
    blk.eval()
    blk.load_state_dict(sd)                 # the reference evaluated AFTER its training forward: use those running stats
    blk.batch_norm.running_mean.copy_(torch.from_numpy(g["l0:running_mean_after"]))
    blk.batch_norm.running_var.copy_(torch.from_numpy(g["l0:running_var_after"]))
    with torch.no_grad():
        y = blk(x)
    assert rel_l2(y, g["l0:y_eval"]) < 1.5e-2


def test_ctc_module_matches_torch(pkg):
    from wav2letter_pytorch_b200.ctc_loss import CTCLoss
    gen = torch.Generator().manual_seed(11)
    N, T, C, S = 8, 300, 29, 60
    out = torch.log_softmax(torch.randn(N, T, C, generator=gen), -1)
    tg = torch.randint(1, C, (N, S), generator=gen, dtype=torch.int32)
    il = torch.randint(200, T + 1, (N,), generator=gen, dtype=torch.int32)
    tl = torch.randint(10, S + 1, (N,), generator=gen, dtype=torch.int32)
    ref_in = out.double().clone().requires_grad_(True)
    ref = torch.nn.CTCLoss(blank=0, reduction="mean", zero_infinity=True)(ref_in.transpose(0, 1), tg, il, tl)
    (3 * ref).backward()
    mine_in = out.cuda().requires_grad_(True)
    crit = CTCLoss(blank=0, reduction="mean", zero_infinity=True)
    loss = crit(mine_in.transpose(0, 1), tg.cuda(), il.cuda(), tl.cuda())       # the reference's exact call shape
    (3 * loss).backward()
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item())
    assert (mine_in.grad.cpu().double() - ref_in.grad).abs().max() <= 2e-3 * ref_in.grad.abs().max()
    # int64 lengths and CPU length tensors behave the same
    loss2 = crit(mine_in.detach().transpose(0, 1), tg.long().cuda(), il.long(), tl.long())
    assert abs(loss2.item() - loss.item()) < 1e-6
    with pytest.raises(RuntimeError):
        crit(out.transpose(0, 1), tg, il, tl)                                  # CPU tensors: no fallback


def test_decoder_interface(pkg, golden):
    from wav2letter_pytorch_b200.decoder import GreedyDecoder
    g = golden("decoder")
    for name in sorted({k.split(":")[0] for k in g.files}):
        labels = [str(s) for s in g[name + ":labels"]]
        sizes = g[name + ":sizes"]
        sizes = None if sizes[0] == -1 else [int(s) for s in sizes]
        probs = torch.from_numpy(g[name + ":probs"])
        strings, offsets = GreedyDecoder(labels).decode(probs.cuda(), sizes, return_offsets=True)
        assert strings == [str(s) for s in g[name + ":strings"]], name
        for i, o in enumerate(offsets):
            assert o[0].dtype == torch.int32 and o[0].tolist() == list(g["%s:offsets:%d" % (name, i)]), name
    # reference unit test (unit_tests/decoder_test.py:40-42): CPU tensor in, 2-D input promoted, sizes=None
    dec = GreedyDecoder(["_", "A", "B", " "], blank_index=0)
    assert dec.decode(torch.FloatTensor([[0.8, 0.2, 0, 0], [0.6, 0.4, 0, 0]]).unsqueeze(0), sizes=None) == [""]
    assert dec.decode(torch.FloatTensor([[0.1, 0.9, 0, 0], [0.6, 0.4, 0, 0]])) == ["A"]
    assert dec.cer_ratio("ab cd", "ab d") == (1, 4) and dec.wer_ratio("ab cd", "ab d") == (1, 2)


def test_prefix_beam_search_on_device_scores(pkg, golden):
    """SURVEY 8f-4 behind the Decoder interface on the box that has the GPU: probabilities produced by the device (Jasper's eval-mode
    softmax head, jasper.py:470-473) go through PrefixBeamSearchLMDecoder.decode as CUDA tensors; transcripts must equal the oracle's
    restatement of decoder.py:147-231 on the same numbers, and the reference's golden cases must come out bit for bit (string AND
    float64 score) from the library routine"""
    from wav2letter_pytorch_b200 import config
    from wav2letter_pytorch_b200.decoder import PrefixBeamSearchLMDecoder, prefix_beam_search
    from wav2letter_pytorch_b200.jasper import Jasper
    g = golden("beam")
    labels = [str(s) for s in g["labels"]]
    for name in ("sharp", "flat", "k1", "beta0", "betaf", "f32"):
        k, beta, prune = g[name + ":params"]
        beta = int(beta) if float(beta).is_integer() else float(beta)
        string, score = prefix_beam_search(g[name + ":probs"], labels, 0, None, int(k), 0.3, beta, float(prune), return_weights=True)
        assert string == str(g[name + ":string"]) and score == float(g[name + ":score"]), name
    torch.manual_seed(3)
    cfg = config.compose(overrides=["model=jasper", "model.mid_layers=3"]).model
    model = Jasper(cfg).cuda().eval()
    x, il, _, _ = O.synthetic_batch(3, 1, seed=4)
    with torch.no_grad():
        probs, _ = model(x.cuda(), il.cuda())                 # [3, T', 29] probabilities on the device
    assert probs.is_cuda and abs(float(probs[0, 0].sum()) - 1.0) < 1e-4
    sharp = torch.softmax(torch.log(probs.clamp_min(1e-30)) * 6.0, -1)          # peaked, so that the beams differ from greedy
    dec = PrefixBeamSearchLMDecoder(None, model.labels, blank_index=0, k=5, alpha=0.3, beta=5, prune=1e-3)
    for p in (probs, sharp):
        got = dec.decode(p)
        want = [O.prefix_beam_search(u.double().cpu().numpy(), list(model.labels), 0, None, 5, 0.3, 5, 1e-3)[0] for u in p]
        assert got == want


def test_bn_scratch_buffers_are_recycled_correctly(pkg):
    """the per-layer BatchNorm scratch (layers.BnScratch: forward statistics / backward sums, each cleared by a kernel of the other
    pass) under call orders a training loop does not produce: a train-mode forward without backward, two forwards before their
    backwards, a second backward over a retained graph -- gradients and running statistics must come out as in plain steps"""
    from wav2letter_pytorch_b200 import config
    from wav2letter_pytorch_b200.wav2letter import Wav2Letter
    cfg = config.compose(overrides=["model.mid_layers=2"]).model
    for l in cfg.layers:
        l["dropout"] = 0.0
    torch.manual_seed(1)
    model = Wav2Letter(cfg).cuda().train()
    xa, il, tg, tl = O.synthetic_batch(3, 1, seed=7)
    xb = O.synthetic_batch(3, 1, seed=8)[0]
    xa, xb, il, tg, tl = xa.cuda(), xb.cuda(), il.cuda(), tg.cuda(), tl.cuda()

    def loss_of(x):
        out, ol = model(x, il)
        return model.criterion(out.transpose(0, 1), tg, ol, tl)

    def grads_of(x):
        model.zero_grad(set_to_none=True)
        loss_of(x).backward()
        return [p.grad.detach().clone() for p in model.parameters()]

    def same(a, b):
        return all(rel_l2(u, v) < 2e-3 or float(v.abs().max()) == 0.0 for u, v in zip(a, b))     # fp32 atomics order only

    ga, gb = grads_of(xa), grads_of(xb)
    assert same(grads_of(xa), ga)                                 # steady state: forward, backward, forward, backward
    with torch.no_grad():                                         # a train-mode forward that never gets its backward (dirty statistics)
        loss_of(xb)
    assert same(grads_of(xa), ga)
    model.zero_grad(set_to_none=True)                             # two forwards, then their backwards in reverse and in order
    la, lb = loss_of(xa), loss_of(xb)
    lb.backward()
    g2 = [p.grad.detach().clone() for p in model.parameters()]
    model.zero_grad(set_to_none=True)
    la.backward()
    assert same(g2, gb) and same([p.grad for p in model.parameters()], ga)
    model.zero_grad(set_to_none=True)                             # a second backward over a retained graph
    l = loss_of(xa)
    l.backward(retain_graph=True)
    model.zero_grad(set_to_none=True)
    l.backward()
    assert same([p.grad for p in model.parameters()], ga)
    rm = model.conv1ds.conv1d_0.batch_norm.running_mean
    assert torch.isfinite(rm).all() and int(model.conv1ds.conv1d_0.batch_norm.num_batches_tracked) == 8


def test_novograd_golden(pkg, golden):
    from wav2letter_pytorch_b200.novograd import Novograd
    g = golden("novograd")
    p = [torch.nn.Parameter(torch.from_numpy(g["p0:0"]).cuda()), torch.nn.Parameter(torch.from_numpy(g["p0:1"]).cuda())]
    opt = Novograd(p, lr=0.01, betas=(0.95, 0.5), weight_decay=1e-3)
    for step in range(3):
        for i, q in enumerate(p):
            q.grad = torch.from_numpy(g["g%d:%d" % (step, i)]).cuda()
        opt.step()
        for i, q in enumerate(p):
            np.testing.assert_allclose(q.detach().cpu().numpy(), g["p%d:%d" % (step + 1, i)], rtol=1e-5, atol=1e-6)
    assert opt.state[p[0]]["exp_avg_sq"].dim() == 0 and opt.state[p[0]]["step"] == 3
    # amsgrad=True (novograd.py:98-102)
    p = [torch.nn.Parameter(torch.from_numpy(g["p0:0"]).cuda()), torch.nn.Parameter(torch.from_numpy(g["p0:1"]).cuda())]
    opt = Novograd(p, lr=0.01, betas=(0.95, 0.5), weight_decay=1e-3, amsgrad=True)
    for step in range(3):
        for i, q in enumerate(p):
            q.grad = torch.from_numpy(g["g%d:%d" % (step, i)]).cuda() * (0.5 ** step)
        opt.step()
        for i, q in enumerate(p):
            np.testing.assert_allclose(q.detach().cpu().numpy(), g["ams_p%d:%d" % (step + 1, i)], rtol=1e-5, atol=1e-6)
    assert float(opt.state[p[0]]["max_exp_avg_sq"]) >= float(opt.state[p[0]]["exp_avg_sq"])


def test_novograd_state_reload(pkg, golden):
    """load_state_dict after a step must drop the cached fused-step plan (it aliases the OLD moments), and moments that arrive in
    another memory layout than the parameter's (a checkpoint of the reference: contiguous [Cout,Cin,k]) are re-laid out"""
    import copy
    from wav2letter_pytorch_b200.novograd import Novograd
    g = golden("novograd")
    # parameter 0 is a permuted view over other-order storage, like the conv weights of this package
    base = torch.from_numpy(g["p0:0"]).cuda()
    perm = list(range(base.dim()))[::-1]
    store = base.permute(*perm).contiguous()
    p = [torch.nn.Parameter(store.permute(*perm)), torch.nn.Parameter(torch.from_numpy(g["p0:1"]).cuda())]
    assert p[0].shape == base.shape and (base.dim() < 2 or not p[0].is_contiguous())
    opt = Novograd(p, lr=0.01, betas=(0.95, 0.5), weight_decay=1e-3)

    def run(step):
        for i, q in enumerate(p):
            q.grad = torch.from_numpy(g["g%d:%d" % (step, i)]).cuda()
        opt.step()

    run(0)
    saved = copy.deepcopy(opt.state_dict())
    for st in saved["state"].values():                       # what a checkpoint of the reference holds: contiguous moments
        st["exp_avg"] = st["exp_avg"].contiguous().cpu()
        st["exp_avg_sq"] = st["exp_avg_sq"].clone().cpu()
    p1 = [q.detach().clone() for q in p]
    run(1)
    run(2)
    with torch.no_grad():
        for q, q1 in zip(p, p1):
            q.copy_(q1)
    opt.load_state_dict(saved)
    run(1)
    run(2)
    for i, q in enumerate(p):
        np.testing.assert_allclose(q.detach().cpu().numpy(), g["p3:%d" % i], rtol=1e-5, atol=1e-6)
    assert opt.state[p[0]]["exp_avg"].stride() == p[0].stride()
    # a gradient set that alternates (parameter 1 without a gradient for one step) must not revive a stale plan either
    p[1].grad = None
    p[0].grad = torch.ones_like(p[0])
    opt.step()
    v_before = float(opt.state[p[1]]["exp_avg_sq"])
    run(2)
    assert float(opt.state[p[1]]["exp_avg_sq"]) != v_before or v_before == 0.0
    assert all(torch.isfinite(q).all() for q in p)


def test_training_step_end_to_end(pkg):
    """training_step on a synthetic collated batch (data_loader.py:149-158 layout) with fused NovoGrad: loss decreases,
    weights and bf16 shadows move together, and the step matches the CPU oracle's first loss within the bf16 tolerance."""
    from wav2letter_pytorch_b200 import config
    from wav2letter_pytorch_b200.wav2letter import Wav2Letter
    cfg = config.compose(overrides=["model.mid_layers=2", "optimizer=novograd"]).model
    for l in cfg.layers:
        l["dropout"] = 0.0
    cfg.optimizer["lr"] = 0.02
    torch.manual_seed(0)
    model = Wav2Letter(cfg).cuda().train()
    (opt,), _ = model.configure_optimizers()
    x, il, tg, tl = O.synthetic_batch(4, 2, seed=3)
    texts = ["".join(O.ENGLISH_LOWERCASE[c] for c in row.tolist()) for row in tg]
    batch = (x.cuda(), il.cuda(), tg.cuda(), tl.cuda(), None, texts)
    sd0 = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    lp_ref, ol_ref = O.w2l_forward(x, il, sd0, O.w2l_layer_specs(2, dropout=0.0), training=True, update_running=False)
    loss_ref = torch.nn.CTCLoss(blank=0, reduction="mean", zero_infinity=True)(lp_ref.transpose(0, 1), tg, ol_ref, tl)
    losses = []
    for it in range(8):
        opt.zero_grad()
        loss = model.training_step(batch, it)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert abs(losses[0] - loss_ref.item()) < 2e-2 * abs(loss_ref.item())
    assert losses[-1] < losses[0]
    assert set(model.logged) == {"train_loss", "learning_rate", "train_cer", "train_wer", "train_len_ratio"}
    conv = model.conv1ds.conv1d_1.conv1
    assert torch.equal(conv.packed().float(), conv.storage().to(torch.bfloat16).float())


@pytest.mark.parametrize("fixture", ["jasper_dense", "jasper_small"])
def test_jasper_dense_golden(pkg, golden, fixture):
    """Jasper with masks, stride-2 prologue, repeats, residual 1x1+BN branches, dilation, unmasked head, softmax in eval;
    ``jasper_small`` additionally has a separable (depthwise + pointwise) block as in the shipped model/jasper.yaml."""
    check_jasper_golden(pkg, golden(fixture), seed=4 if fixture == "jasper_dense" else 2)


def test_jasper_odd_widths_golden(pkg, golden):
    """Jasper at channel counts that are padded internally (161 STFT bins in; widths 100 / 36 / 250 / 52 over a stride-2 prologue, a
    residual dense block, a separable block and a dilated one): same checks and bounds as the other Jasper fixtures"""
    check_jasper_golden(pkg, golden("jasper_odd"), seed=23, input_size=0)


def check_jasper_golden(pkg, g, seed, input_size=None):
    from wav2letter_pytorch_b200 import config
    from wav2letter_pytorch_b200.jasper import Jasper
    blocks = [dict(b, dropout=0) for b in json.loads(str(g["blocks_json"]))]
    cfg = config.compose(overrides=["model=jasper", "model.mid_layers=%d" % len(blocks)]).model
    cfg["jasper_blocks"] = config.to_attr(blocks)
    if input_size is not None:
        cfg["input_size"] = input_size
    torch.manual_seed(seed)
    model = Jasper(cfg)
    for k in g.files:                                                   # seeded construction == the reference's
        if k.startswith("sd_init:"):
            assert np.array_equal(model.state_dict()[k[8:]].numpy(), g[k]), k
    assert sorted(model.state_dict().keys()) == sorted(k[4:] for k in g.files if k.startswith("sd0:"))
    _load_sd(model, g, "sd0:")
    model.cuda().train()
    x, il = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["il"]).cuda()
    tg, tl = torch.from_numpy(g["tg"]).cuda(), torch.from_numpy(g["tl"]).cuda()
    hs, taps, rows, out, ol, loss = L.jasper_run_blocks(model, x, il, tg, tl)       # ONE training forward through Jasper.forward
    assert np.array_equal(ol.cpu().numpy(), g["train:out_len"]) and ol.dtype == torch.int64
    assert rel_l2(out.detach(), g["train:out"]) < 2e-2
    assert abs(loss.item() - float(g["train:loss"])) < 2e-2 * abs(float(g["train:loss"]))
    # gradients, closed tolerances: every conv+BN group on its own (see check_w2l_golden), then one fixed end-to-end bound
    table = L.jasper_layerwise_table(model, O.jasper_block_specs(blocks), hs, taps, rows, out)
    bad = L.check_table(table, TOL_EMU, TOL_REF)
    assert not bad, (bad, L.format_table(table))
    for name, p in model.named_parameters():
        ref = torch.from_numpy(g["train:grad:" + name])
        assert p.grad is not None and p.grad.shape == p.shape, name
        err_ref = rel_l2(p.grad, ref)
        assert err_ref < E2E_TOY_BOUND, (name, err_ref)
    for k in g.files:
        if k.startswith("sd1:") and "running" in k:
            np.testing.assert_allclose(model.state_dict()[k[4:]].cpu().numpy(), g[k], rtol=2e-2, atol=2e-3, err_msg=k)
    model.eval()
    with torch.no_grad():
        o, ol = model(x, il)
    assert rel_l2(o, g["eval:out"]) < 2e-2 and abs(float(o.sum(-1).mean()) - 1.0) < 1e-5    # probabilities in eval


def test_jasper_dense_golden_fp32_faithful_mode(pkg, golden):
    """the dense Jasper fixture (masks, stride-2 prologue, repeats, residual 1x1+BN branches, dilation) with precision='tf32': fp32
    storage end to end incl. the residual variants of the BatchNorm passes and zero-padded (negative row offset) tf32 weight gradients"""
    from wav2letter_pytorch_b200 import config
    from wav2letter_pytorch_b200.jasper import Jasper
    g = golden("jasper_dense")
    blocks = [dict(b, dropout=0) for b in json.loads(str(g["blocks_json"]))]
    cfg = config.compose(overrides=["model=jasper", "model.mid_layers=%d" % len(blocks)]).model
    cfg["jasper_blocks"] = config.to_attr(blocks)
    cfg["precision"] = "tf32"
    torch.manual_seed(4)
    model = Jasper(cfg)
    _load_sd(model, g, "sd0:")
    model.cuda().train()
    x, il = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["il"]).cuda()
    tg, tl = torch.from_numpy(g["tg"]).cuda(), torch.from_numpy(g["tl"]).cuda()
    out, ol = model(x, il)
    loss = model.criterion(out.transpose(0, 1), tg, ol, tl)
    loss.backward()
    assert np.array_equal(ol.cpu().numpy(), g["train:out_len"])
    report = {"out": rel_l2(out.detach(), g["train:out"]), "loss": abs(loss.item() - float(g["train:loss"])) / abs(float(g["train:loss"]))}
    worst = ("", 0.0)
    for name, p in model.named_parameters():
        assert p.grad is not None and p.grad.dtype == torch.float32, name
        e = rel_l2(p.grad, torch.from_numpy(g["train:grad:" + name]))
        if e > worst[1]:
            worst = (name, e)
    report["grad"] = worst[1]
    run = 0.0
    for k in g.files:
        if k.startswith("sd1:") and "running" in k:
            run = max(run, rel_l2(model.state_dict()[k[4:]], g[k]))
    report["running"] = run
    print("Jasper tf32 mode vs the reference fixture:", json.dumps(report), "worst gradient:", worst[0])
    # (the worst gradient here is a BatchNorm bias deep inside a residual block -- a sum that cancels to ~1 % of its terms -- at
    # 0.12 on the host emulation, where the bf16 path is allowed 0.35; every weight gradient is below 0.06)
    for key, bound in dict(TOL_TF32, grad=2e-1).items():
        assert report[key] < bound, (key, report, worst)
    sd1 = {k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd1:")}
    model.load_state_dict(sd1, strict=False)      # the reference's own running statistics after its step (the fixture holds only those)
    model.eval()
    with torch.no_grad():
        o, _ = model(x, il)
    # eval probabilities: 1.9e-3 on the host emulation (running statistics do not re-centre the tf32 roundings of 11 convs the way
    # batch statistics do in training: 4.2e-4 there); bf16 path: 2e-2
    assert rel_l2(o, g["eval:out"]) < 5e-3 and abs(float(o.sum(-1).mean()) - 1.0) < 1e-5


@pytest.mark.parametrize("fixture", ["jasper_small", "jasper_odd"])
def test_jasper_separable_golden_fp32_faithful_mode(pkg, golden, fixture):
    """``jasper_small`` -- with a separable (depthwise + pointwise) block as in the shipped model/jasper.yaml -- in precision='tf32': the
    depthwise convs run as plain fp32 FMAs over fp32 activations (w2l_depthwise_*_f32); ``jasper_odd``: the same mode over internally
    padded channel counts"""
    from wav2letter_pytorch_b200 import config
    from wav2letter_pytorch_b200.jasper import Jasper
    g = golden(fixture)
    blocks = [dict(b, dropout=0) for b in json.loads(str(g["blocks_json"]))]
    assert any(b.get("separable", True) for b in blocks)
    cfg = config.compose(overrides=["model=jasper", "model.mid_layers=%d" % len(blocks)]).model
    cfg["jasper_blocks"] = config.to_attr(blocks)
    cfg["precision"] = "tf32"
    if fixture == "jasper_odd":
        cfg["input_size"] = 0
    torch.manual_seed(2)
    model = Jasper(cfg)
    _load_sd(model, g, "sd0:")
    model.cuda().train()
    x, il = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["il"]).cuda()
    tg, tl = torch.from_numpy(g["tg"]).cuda(), torch.from_numpy(g["tl"]).cuda()
    out, ol = model(x, il)
    loss = model.criterion(out.transpose(0, 1), tg, ol, tl)
    loss.backward()
    assert np.array_equal(ol.cpu().numpy(), g["train:out_len"])
    report = {"out": rel_l2(out.detach(), g["train:out"]), "loss": abs(loss.item() - float(g["train:loss"])) / abs(float(g["train:loss"]))}
    worst = ("", 0.0)
    for name, p in model.named_parameters():
        assert p.grad is not None and p.grad.dtype == torch.float32, name
        e = rel_l2(p.grad, torch.from_numpy(g["train:grad:" + name]))
        if e > worst[1]:
            worst = (name, e)
    report["grad"] = worst[1]
    print("separable Jasper tf32 mode vs the reference fixture:", json.dumps(report), "worst gradient:", worst[0])
    for key, bound in dict(out=TOL_TF32["out"], loss=TOL_TF32["loss"], grad=2e-1).items():
        assert report[key] < bound, (key, report, worst)


@pytest.mark.parametrize("mid_layers", [1, 20])
def test_baseline_config1_forward_ctc_decode(pkg, mid_layers):
    """BASELINE.json configs[0]: Wav2Letter default config, forward + CTCLoss + greedy decode, batch 8, synthetic 10 s utterances
    (T=1001 frames -> T'=500), English labels -- the CUDA path against the CPU oracle (fp32 torch) on the same weights and batch.
    Stated tolerances: log-probs rel-L2 <= 1e-2 for the literal default (mid_layers=1) and <= 3e-2 through the 20-block stack
    (bf16 operands / bf16-stored activations, fp32 accumulation); CTC loss <= 1e-4 relative on identical log-probs; transcripts
    and offsets bit-exact on identical scores."""
    from wav2letter_pytorch_b200 import config
    from wav2letter_pytorch_b200.wav2letter import Wav2Letter
    cfg = config.compose(overrides=["model.mid_layers=%d" % mid_layers]).model
    torch.manual_seed(3)
    model = Wav2Letter(cfg)
    # give the BatchNorm buffers non-trivial values, as after training
    g = torch.Generator().manual_seed(4)
    for n, b in model.named_buffers():
        if n.endswith("running_mean"):
            b.copy_(0.1 * torch.randn(b.shape, generator=g))
        if n.endswith("running_var"):
            b.copy_(0.5 + torch.rand(b.shape, generator=g))
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model.cuda().eval()
    x, il, tg, tl = O.synthetic_batch(8, 10, seed=2, ragged=True)
    assert x.shape == (8, 64, 1001)
    with torch.no_grad():
        out, ol = model(x.cuda(), il.cuda())
    specs = O.w2l_layer_specs(mid_layers)
    ref, ref_ol = O.w2l_forward(x, il, sd, specs, training=False)
    assert out.shape == ref.shape == (8, 500, 29) and np.array_equal(ol.cpu().numpy(), ref_ol.numpy())
    assert rel_l2(out, ref) < (1e-2 if mid_layers == 1 else 3e-2), rel_l2(out, ref)
    # CTC loss of the reference's criterion vs ours on IDENTICAL log-probs
    crit = torch.nn.CTCLoss(blank=0, reduction="mean", zero_infinity=True)
    want = crit(out.detach().cpu().transpose(0, 1), tg, ol.cpu(), tl)
    got = model.criterion(out.transpose(0, 1), tg.cuda(), ol, tl.cuda())
    assert abs(got.item() - want.item()) <= 1e-4 * abs(want.item()), (got.item(), want.item())
    # greedy transcripts on identical scores: the reference's per-frame loop (oracle) vs the CUDA argmax+collapse
    hyp, offs = model.ctc_decoder.decode(out, ol, return_offsets=True)
    ref_hyp, ref_offs = O.greedy_decode(out.cpu().numpy(), ol.cpu().numpy())
    assert hyp == ref_hyp
    assert [o[0].tolist() for o in offs] == [list(o) for o in ref_offs]


def test_jasper_nan_assert(pkg, golden):
    """jasper.py:474 `assert not (jasper_res != jasper_res).any()`: AssertionError on NaN scores, in each of the three modes"""
    from wav2letter_pytorch_b200 import config
    from wav2letter_pytorch_b200.jasper import Jasper
    g = golden("jasper_dense")
    blocks = [dict(b, dropout=0) for b in json.loads(str(g["blocks_json"]))]
    cfg = config.compose(overrides=["model=jasper", "model.mid_layers=5"]).model
    cfg["jasper_blocks"] = config.to_attr(blocks)
    torch.manual_seed(4)
    model = Jasper(cfg).cuda().eval()
    x, il = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["il"]).cuda()
    with torch.no_grad():
        model(x, il)                                         # clean input: no assertion
        bad = x.clone()
        bad[0, 3, 5] = float("nan")
        with pytest.raises(AssertionError):
            model(bad, il)                                   # default: checked at once, like the reference
        model.nan_check = "deferred"
        model(bad, il)                                       # no sync here ...
        with pytest.raises(AssertionError):
            model.check_nan()                                # ... raised when the flag is looked at
        model(x, il)
        model.check_nan()
        model.nan_check = "off"
        model(bad, il)


def test_gradient_accumulation_and_zero_grad_in_place(pkg):
    """two backward passes into existing .grad tensors (gradient accumulation, zero_grad(set_to_none=False)): autograd adds the new
    weight gradient on the compute stream right after backward() returns, so those wgrads must not run on the side stream"""
    from wav2letter_pytorch_b200 import config, layers
    from wav2letter_pytorch_b200.wav2letter import Wav2Letter
    cfg = config.compose(overrides=["model.mid_layers=3"]).model
    for l in cfg.layers:
        l["dropout"] = 0.0
    torch.manual_seed(0)
    model = Wav2Letter(cfg).cuda().train()
    x, il, tg, tl = (t.cuda() for t in O.synthetic_batch(6, 4, seed=5))

    def backward_once():
        out, ol = model(x, il)
        model.criterion(out.transpose(0, 1), tg, ol, tl).backward()

    model.zero_grad(set_to_none=True)
    backward_once()
    torch.cuda.synchronize()
    single = [p.grad.detach().clone() for p in model.parameters()]
    backward_once()                                            # accumulates into the existing tensors
    torch.cuda.synchronize()
    for p, g1 in zip(model.parameters(), single):
        if g1.abs().max() > 0:
            assert rel_l2(p.grad, 2 * g1) < 2e-2
    model.zero_grad(set_to_none=False)                         # grads stay allocated (zeros): the next backward adds in place
    backward_once()
    torch.cuda.synchronize()
    for p, g1 in zip(model.parameters(), single):
        if g1.abs().max() > 0:
            assert rel_l2(p.grad, g1) < 2e-2
