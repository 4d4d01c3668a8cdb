"""TEST INFRASTRUCTURE ONLY -- freezes outputs of the UNMODIFIED reference into tests/golden/*.npz.

Run in the build container (needs /root/reference):   python -m oracle.gen_golden
The fixtures are committed; tests read only the .npz files, never /root/reference.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader as rl  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def _sd(model, prefix="sd:"):
    return {prefix + k: v.detach().clone() for k, v in model.state_dict().items()}


def _train_step_record(ref, model, x, il, tg, tl, texts):
    """forward -> criterion exactly as base_asr_models.py:78-81, then backward."""
    model.zero_grad()
    out, ol = model.forward(x, il)
    loss = model.criterion(out.transpose(0, 1), tg, ol, tl)
    loss.backward()
    rec = {"out": out, "out_len": ol, "loss": loss}
    for n, p in model.named_parameters():
        rec["grad:" + n] = p.grad.clone()
    rec["decoded"] = np.array(model.ctc_decoder.decode(out, ol))
    return rec


def gen_decoder(ref):
    D = ref.decoder.GreedyDecoder
    cases = {}
    # unit_tests/decoder_test.py:40-42
    lab = ["_", "A", "B", " "]
    p = torch.FloatTensor([[0.8, 0.2, 0, 0], [0.6, 0.4, 0, 0]]).unsqueeze(0)
    cases["unit"] = (lab, p, None)
    # decoder.py:305-311 (__main__ vectors)
    lab2 = ["_", "a", "b", " "]
    cases["main_a"] = (lab2, torch.Tensor([[[0.4, 0.6, 0, 0]]]), None)
    cases["main_space"] = (lab2, torch.Tensor([[[0.4, 0.1, 0, 0.5]]]), None)
    cases["main_aba"] = (lab2, torch.Tensor([[[0.0, 0.6, 0.3, 0.1], [0.0, 0.6, 0.3, 0.1], [0.0, 0.3, 0.6, 0.1],
                                             [0.0, 0.6, 0.3, 0.1]], [[0.4, 0.1, 0, 0.5]] * 4]), [4, 1])
    # ties, NaN, a _ a, sizes truncation, random
    cases["ties"] = (lab2, torch.Tensor([[[0.5, 0.5, 0, 0], [0.1, 0.4, 0.4, 0.1], [0, 0, 0.5, 0.5]]]), None)
    cases["a_blank_a"] = (lab2, torch.Tensor([[[0, 1, 0, 0], [1, 0, 0, 0], [0, 1, 0, 0], [0, 1, 0, 0]]]), None)
    nan = float("nan")
    cases["nan"] = (lab2, torch.Tensor([[[0.9, nan, 0.1, 0], [0.2, 0.1, nan, nan], [nan, 0.9, 0, 0], [0, 0, 0, 1]]]), None)
    g = torch.Generator().manual_seed(3)
    eng = list(ref.label_sets.labels_map["english_lowercase"])
    pr = torch.softmax(3 * torch.randn(5, 97, 29, generator=g), -1)
    pr[:, :, 0] += 0.15
    cases["random"] = (eng, pr, [97, 50, 1, 0, 96])
    lpq = torch.round(torch.log_softmax(torch.randn(3, 64, 29, generator=g), -1) * 2) / 2   # many exact ties
    cases["quantised"] = (eng, lpq, None)
    out = {}
    for name, (lab, p, sizes) in cases.items():
        strings, offsets = D(lab).decode(p, sizes=sizes, return_offsets=True)
        out[name + ":labels"] = np.array(lab)
        out[name + ":probs"] = p.numpy()
        out[name + ":sizes"] = np.array(sizes if sizes is not None else [-1])
        out[name + ":strings"] = np.array(strings)
        for i, o in enumerate(offsets):
            out[name + ":offsets:%d" % i] = o[0].numpy()
        _, am = torch.max(p, 2)
        out[name + ":argmax"] = am.numpy()
    np.savez_compressed(os.path.join(OUT, "decoder.npz"), **out)


def gen_conv_block(ref):
    torch.manual_seed(0)
    out = {}
    for tag, args in {"l0": (64, 256, (11,), 2, 0.0, 1), "dil": (32, 48, (5,), 1, -1.0, 2),
                      "odd": (33, 16, (4,), 2, -1.0, 1)}.items():
        blk = ref.wav2letter.Conv1dBlock(*args[:4], drop_out_prob=args[4], dilation=args[5])
        with torch.no_grad():
            blk.batch_norm.weight.uniform_(0.5, 1.5)
            blk.batch_norm.bias.uniform_(-0.5, 0.5)
        x = torch.randn(2, args[0], 61)
        out.update(_np({tag + ":" + k: v for k, v in _sd(blk, "").items()}))
        out[tag + ":x"] = x.numpy()
        blk.train()
        out[tag + ":y_train"] = blk(x).detach().numpy()
        out[tag + ":running_mean_after"] = blk.batch_norm.running_mean.numpy().copy()
        out[tag + ":running_var_after"] = blk.batch_norm.running_var.numpy().copy()
        blk.eval()
        out[tag + ":y_eval"] = blk(x).detach().numpy()
        out[tag + ":pad"] = np.array(blk.paddingAdded.padding if hasattr(blk.paddingAdded, "padding") else (0, 0))
    np.savez_compressed(os.path.join(OUT, "conv_block.npz"), **out)


def gen_w2l(ref):
    torch.manual_seed(1)
    cfg = rl.reference_model_cfg("wav2letter", mid_layers=3, dropout=-1)
    small = [dict(output_size=64, kernel_size=11, stride=2, dilation=1, dropout=-1),
             dict(output_size=128, kernel_size=13, stride=1, dilation=1, dropout=-1),
             dict(output_size=64, kernel_size=5, stride=1, dilation=2, dropout=-1)]
    cfg["layers"] = rl.to_attr(small)
    model = ref.wav2letter.Wav2Letter(cfg)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.2, 0.2)
    x = torch.randn(3, 64, 201)
    il = torch.tensor([201, 160, 121], dtype=torch.int32)
    tg = torch.randint(1, 29, (3, 20), dtype=torch.int32)
    tl = torch.tensor([20, 12, 7], dtype=torch.int32)
    for n in range(3):
        tg[n, tl[n]:] = 0
    out = {"x": x, "il": il, "tg": tg, "tl": tl, "layers": np.array([[l["output_size"], l["kernel_size"], l["stride"],
                                                                     l["dilation"]] for l in small])}
    out.update(_sd(model, "sd0:"))
    model.train()
    rec = _train_step_record(ref, model, x, il, tg, tl, None)
    out.update({"train:" + k: v for k, v in rec.items()})
    out.update(_sd(model, "sd1:"))                       # running stats after one training forward
    model.eval()
    with torch.no_grad():
        o, ol = model(x, il)
    out["eval:out"], out["eval:out_len"] = o, ol
    out["eval:decoded"] = np.array(model.ctc_decoder.decode(o, ol))
    out["scaling_factor"] = model.scaling_factor
    np.savez_compressed(os.path.join(OUT, "w2l_small.npz"), **_np(out))


def gen_jasper(ref):
    torch.manual_seed(2)
    blocks = [dict(layer_size=64, kernel_size=10, stride=2, residual=False, separable=False, repeat=1),
              dict(layer_size=64, kernel_size=11, stride=1, residual=True, separable=False, repeat=3),
              dict(layer_size=128, kernel_size=12, stride=1, residual=True, separable=True, repeat=2),
              dict(layer_size=128, kernel_size=7, stride=1, dilation=2, residual=False, separable=False, repeat=1),
              dict(layer_size=64, kernel_size=1, stride=1, residual=False, separable=False, repeat=1)]
    cfg = rl.reference_model_cfg("jasper", mid_layers=5, dropout=0, jasper_blocks=blocks)
    model = ref.jasper.Jasper(cfg)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.2, 0.2)
    x = torch.randn(3, 64, 201)
    il = torch.tensor([201, 160, 121], dtype=torch.int32)
    tg = torch.randint(1, 29, (3, 20), dtype=torch.int32)
    tl = torch.tensor([20, 12, 7], dtype=torch.int32)
    for n in range(3):
        tg[n, tl[n]:] = 0
        x[n, :, il[n]:] = 0
    import json
    out = {"x": x, "il": il, "tg": tg, "tl": tl, "blocks_json": np.array(json.dumps(blocks))}
    out.update(_sd(model, "sd0:"))
    model.train()
    rec = _train_step_record(ref, model, x, il, tg, tl, None)
    out.update({"train:" + k: v for k, v in rec.items()})
    out.update(_sd(model, "sd1:"))
    model.eval()
    with torch.no_grad():
        o, ol = model(x, il)
    out["eval:out"], out["eval:out_len"] = o, ol
    out["scaling_factor"] = model.scaling_factor
    np.savez_compressed(os.path.join(OUT, "jasper_small.npz"), **_np(out))


def gen_jasper_dense(ref):
    """dense (non-separable) Jasper with masks, a stride-2 prologue, repeats, residuals and a dilated block --
    the structure of Jasper 10x5 (BASELINE config 3) at toy widths."""
    torch.manual_seed(4)
    import json
    blocks = [dict(layer_size=64, kernel_size=10, stride=2, residual=False, separable=False, repeat=1),
              dict(layer_size=64, kernel_size=5, stride=1, residual=True, separable=False, repeat=3),
              dict(layer_size=128, kernel_size=6, stride=1, residual=True, separable=False, repeat=2),
              dict(layer_size=128, kernel_size=3, stride=1, dilation=2, residual=False, separable=False, repeat=1),
              dict(layer_size=64, kernel_size=1, stride=1, residual=False, separable=False, repeat=1)]
    cfg = rl.reference_model_cfg("jasper", mid_layers=5, dropout=0, jasper_blocks=blocks)
    model = ref.jasper.Jasper(cfg)
    out = {"blocks_json": np.array(json.dumps(blocks))}
    # a few tensors of the freshly constructed model, to check seeded-init parity (torch.manual_seed(4))
    out.update({k: v for k, v in _sd(model, "sd_init:").items()
                if any(t in k for t in ("encoder.0.mconv.0", "encoder.2.res.0.0", "encoder.1.mconv.8", "final_layer"))})
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.2, 0.2)
    x = torch.randn(3, 64, 201)
    il = torch.tensor([201, 160, 121], dtype=torch.int32)
    tg = torch.randint(1, 29, (3, 20), dtype=torch.int32)
    tl = torch.tensor([20, 12, 7], dtype=torch.int32)
    for n in range(3):
        tg[n, tl[n]:] = 0
        x[n, :, il[n]:] = 0
    out.update({"x": x, "il": il, "tg": tg, "tl": tl})
    out.update(_sd(model, "sd0:"))
    model.train()
    rec = _train_step_record(ref, model, x, il, tg, tl, None)
    out.update({"train:" + k: v for k, v in rec.items()})
    out.update({k: v for k, v in _sd(model, "sd1:").items() if "running" in k or "num_batches" in k})
    model.eval()
    with torch.no_grad():
        o, ol = model(x, il)
    out["eval:out"], out["eval:out_len"] = o, ol
    np.savez_compressed(os.path.join(OUT, "jasper_dense.npz"), **_np(out))


def gen_strided(ref):
    """stride > 1 beyond the first layer (not in the shipped yamls, but legal in both model classes): a Wav2Letter with two
    stride-2 blocks, and a Jasper with a strided dense block (repeat 2: every repeat strides, jasper.py:193-208) and a strided
    separable block; strided Jasper blocks carry no residual (the reference's add would fail on the time dimension)."""
    import json
    torch.manual_seed(9)
    cfg = rl.reference_model_cfg("wav2letter", mid_layers=3, dropout=-1)
    small = [dict(output_size=64, kernel_size=11, stride=2, dilation=1, dropout=-1),
             dict(output_size=128, kernel_size=7, stride=2, dilation=1, dropout=-1),
             dict(output_size=64, kernel_size=5, stride=1, dilation=2, dropout=-1)]
    cfg["layers"] = rl.to_attr(small)
    model = ref.wav2letter.Wav2Letter(cfg)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.2, 0.2)
    x = torch.randn(3, 64, 201)
    il = torch.tensor([201, 160, 121], dtype=torch.int32)
    tl = torch.tensor([14, 9, 5], dtype=torch.int32)
    tg = torch.randint(1, 29, (3, 14), dtype=torch.int32)
    for n in range(3):
        tg[n, tl[n]:] = 0
    out = {"x": x, "il": il, "tg": tg, "tl": tl, "layers": np.array([[l["output_size"], l["kernel_size"], l["stride"],
                                                                     l["dilation"]] for l in small])}
    out.update(_sd(model, "sd0:"))
    model.train()
    rec = _train_step_record(ref, model, x, il, tg, tl, None)
    out.update({"train:" + k: v for k, v in rec.items()})
    out.update(_sd(model, "sd1:"))
    model.eval()
    with torch.no_grad():
        o, ol = model(x, il)
    out["eval:out"], out["eval:out_len"] = o, ol
    out["eval:decoded"] = np.array(model.ctc_decoder.decode(o, ol))
    out["scaling_factor"] = model.scaling_factor
    np.savez_compressed(os.path.join(OUT, "w2l_strided.npz"), **_np(out))

    torch.manual_seed(10)
    blocks = [dict(layer_size=64, kernel_size=10, stride=2, residual=False, separable=False, repeat=1),
              dict(layer_size=64, kernel_size=5, stride=2, residual=False, separable=False, repeat=2),
              dict(layer_size=128, kernel_size=6, stride=2, residual=False, separable=True, repeat=1),
              dict(layer_size=64, kernel_size=3, stride=1, residual=True, separable=False, repeat=2)]
    cfg = rl.reference_model_cfg("jasper", mid_layers=4, dropout=0, jasper_blocks=blocks)
    model = ref.jasper.Jasper(cfg)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.2, 0.2)
    x = torch.randn(3, 64, 401)
    il = torch.tensor([401, 320, 241], dtype=torch.int32)
    tl = torch.tensor([10, 7, 4], dtype=torch.int32)
    tg = torch.randint(1, 29, (3, 10), dtype=torch.int32)
    for n in range(3):
        tg[n, tl[n]:] = 0
        x[n, :, il[n]:] = 0
    out = {"x": x, "il": il, "tg": tg, "tl": tl, "blocks_json": np.array(json.dumps(blocks))}
    out.update(_sd(model, "sd0:"))
    model.train()
    rec = _train_step_record(ref, model, x, il, tg, tl, None)
    out.update({"train:" + k: v for k, v in rec.items()})
    out.update({k: v for k, v in _sd(model, "sd1:").items() if "running" in k or "num_batches" in k})
    model.eval()
    with torch.no_grad():
        o, ol = model(x, il)
    out["eval:out"], out["eval:out_len"] = o, ol
    out["scaling_factor"] = model.scaling_factor
    np.savez_compressed(os.path.join(OUT, "jasper_strided.npz"), **_np(out))


def gen_narrow(ref):
    """hidden widths that are multiples of 8 but not of 16 (72, 136): legal in the reference, they make the backward-data GEMM read
    activation-gradient rows narrower than the padded weight tiles"""
    torch.manual_seed(11)
    cfg = rl.reference_model_cfg("wav2letter", mid_layers=3, dropout=-1)
    small = [dict(output_size=72, kernel_size=11, stride=2, dilation=1, dropout=-1),
             dict(output_size=136, kernel_size=7, stride=1, dilation=1, dropout=-1),
             dict(output_size=72, kernel_size=5, stride=1, dilation=2, dropout=-1)]
    cfg["layers"] = rl.to_attr(small)
    model = ref.wav2letter.Wav2Letter(cfg)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.2, 0.2)
    x = torch.randn(3, 64, 201)
    il = torch.tensor([201, 160, 121], dtype=torch.int32)
    tl = torch.tensor([18, 11, 6], dtype=torch.int32)
    tg = torch.randint(1, 29, (3, 18), dtype=torch.int32)
    for n in range(3):
        tg[n, tl[n]:] = 0
    out = {"x": x, "il": il, "tg": tg, "tl": tl, "layers": np.array([[l["output_size"], l["kernel_size"], l["stride"],
                                                                     l["dilation"]] for l in small])}
    out.update(_sd(model, "sd0:"))
    model.train()
    rec = _train_step_record(ref, model, x, il, tg, tl, None)
    out.update({"train:" + k: v for k, v in rec.items()})
    out.update(_sd(model, "sd1:"))
    model.eval()
    with torch.no_grad():
        o, ol = model(x, il)
    out["eval:out"], out["eval:out_len"] = o, ol
    out["eval:decoded"] = np.array(model.ctc_decoder.decode(o, ol))
    out["scaling_factor"] = model.scaling_factor
    np.savez_compressed(os.path.join(OUT, "w2l_narrow.npz"), **_np(out))


def gen_odd(ref):
    """widths the tensor-core path has to PAD internally: the reference accepts any ``output_size`` (wav2letter.py:59-64; 250 is the
    hidden width of the original paper) and, with ``input_size`` unset, feeds the 161 STFT bins of the default audio config
    (wav2letter.py:52-56) -- none a multiple of 8, one below 64"""
    torch.manual_seed(17)
    cfg = rl.reference_model_cfg("wav2letter", mid_layers=3, dropout=-1)
    cfg["input_size"] = 0                                  # -> int(1 + sample_rate * window_size / 2) = 161
    odd = [dict(output_size=250, kernel_size=5, stride=2, dilation=1, dropout=-1),
           dict(output_size=36, kernel_size=7, stride=1, dilation=1, dropout=-1),
           dict(output_size=250, kernel_size=5, stride=1, dilation=2, dropout=-1)]
    cfg["layers"] = rl.to_attr(odd)
    model = ref.wav2letter.Wav2Letter(cfg)
    assert model.input_size == 161
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.2, 0.2)
    x = torch.randn(3, 161, 201)
    il = torch.tensor([201, 160, 121], dtype=torch.int32)
    tl = torch.tensor([18, 11, 6], dtype=torch.int32)
    tg = torch.randint(1, 29, (3, 18), dtype=torch.int32)
    for n in range(3):
        tg[n, tl[n]:] = 0
    out = {"x": x, "il": il, "tg": tg, "tl": tl, "layers": np.array([[l["output_size"], l["kernel_size"], l["stride"],
                                                                     l["dilation"]] for l in odd])}
    out.update(_sd(model, "sd0:"))
    model.train()
    rec = _train_step_record(ref, model, x, il, tg, tl, None)
    out.update({"train:" + k: v for k, v in rec.items()})
    out.update(_sd(model, "sd1:"))
    model.eval()
    with torch.no_grad():
        o, ol = model(x, il)
    out["eval:out"], out["eval:out_len"] = o, ol
    out["eval:decoded"] = np.array(model.ctc_decoder.decode(o, ol))
    out["scaling_factor"] = model.scaling_factor
    np.savez_compressed(os.path.join(OUT, "w2l_odd.npz"), **_np(out))


def gen_jasper_odd(ref):
    """Jasper with channel counts the tensor-core path has to pad internally (the reference takes any ``layer_size``, jasper.py:289-298):
    161 STFT bins in, a stride-2 prologue, a residual dense block, a separable block and a dilated one at widths 100 / 36 / 250 / 52"""
    import json
    torch.manual_seed(23)
    blocks = [dict(layer_size=100, kernel_size=10, stride=2, residual=False, separable=False, repeat=1),
              dict(layer_size=36, kernel_size=5, stride=1, residual=True, separable=False, repeat=2),
              dict(layer_size=250, kernel_size=7, stride=1, residual=True, separable=True, repeat=2),
              dict(layer_size=52, kernel_size=3, stride=1, dilation=2, residual=False, separable=False, repeat=1)]
    cfg = rl.reference_model_cfg("jasper", mid_layers=4, dropout=0, jasper_blocks=blocks)
    cfg["input_size"] = 0                                  # -> 161 STFT bins
    model = ref.jasper.Jasper(cfg)
    assert model.input_size == 161
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.2, 0.2)
    x = torch.randn(3, 161, 201)
    il = torch.tensor([201, 160, 121], dtype=torch.int32)
    tg = torch.randint(1, 29, (3, 20), dtype=torch.int32)
    tl = torch.tensor([20, 12, 7], dtype=torch.int32)
    for n in range(3):
        tg[n, tl[n]:] = 0
        x[n, :, il[n]:] = 0
    out = {"x": x, "il": il, "tg": tg, "tl": tl, "blocks_json": np.array(json.dumps(blocks))}
    out.update(_sd(model, "sd0:"))
    model.train()
    rec = _train_step_record(ref, model, x, il, tg, tl, None)
    out.update({"train:" + k: v for k, v in rec.items()})
    out.update(_sd(model, "sd1:"))
    model.eval()
    with torch.no_grad():
        o, ol = model(x, il)
    out["eval:out"], out["eval:out_len"] = o, ol
    out["scaling_factor"] = model.scaling_factor
    np.savez_compressed(os.path.join(OUT, "jasper_odd.npz"), **_np(out))


def gen_ctc(ref):
    g = torch.Generator().manual_seed(5)
    crit = torch.nn.CTCLoss(blank=0, reduction="mean", zero_infinity=True)    # base_asr_models.py:23
    out = {}
    cases = {
        "ragged": (4, 50, 29, [50, 41, 30, 17], [12, 9, 0, 8]),
        "infeasible": (3, 6, 5, [6, 2, 4], [3, 2, 4]),       # n=1: T=2,S=2 w/ repeat forced below; n=2: S=T=4 w/ repeats
        "long": (2, 300, 29, [300, 257], [120, 64]),
        "single": (1, 1, 4, [1], [1]),
    }
    for name, (N, T, C, il, tl) in cases.items():
        lp = torch.log_softmax(torch.randn(N, T, C, generator=g), -1)
        S = max(max(tl), 1)
        tg = torch.randint(1, C, (N, S), generator=g, dtype=torch.int32)
        if name == "infeasible":
            tg[1, :2] = torch.tensor([2, 2])
            tg[2, :4] = torch.tensor([1, 1, 3, 3])
        for n in range(N):
            tg[n, tl[n]:] = 0
        lp.requires_grad_(True)
        ilt, tlt = torch.tensor(il, dtype=torch.int32), torch.tensor(tl, dtype=torch.int32)
        loss = crit(lp.transpose(0, 1), tg, ilt, tlt)
        loss.backward()
        nll = torch.nn.functional.ctc_loss(lp.detach().transpose(0, 1), tg, ilt, tlt, blank=0, reduction="none",
                                           zero_infinity=True)
        out.update({name + ":lp": lp.detach(), name + ":tg": tg, name + ":il": ilt, name + ":tl": tlt,
                    name + ":loss": loss.detach(), name + ":grad": lp.grad, name + ":nll": nll})
    np.savez_compressed(os.path.join(OUT, "ctc.npz"), **_np(out))


def gen_novograd(ref):
    torch.manual_seed(7)
    import warnings
    out = {}
    p = [torch.nn.Parameter(torch.randn(16, 8, 3)), torch.nn.Parameter(torch.randn(16))]
    out["p0:0"], out["p0:1"] = p[0].detach().clone(), p[1].detach().clone()
    opt = ref.novograd.Novograd(p, lr=0.01, betas=(0.95, 0.5), weight_decay=1e-3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for step in range(3):
            for i, q in enumerate(p):
                q.grad = torch.randn_like(q)
                out["g%d:%d" % (step, i)] = q.grad.clone()
            opt.step()
            for i, q in enumerate(p):
                out["p%d:%d" % (step + 1, i)] = q.detach().clone()
    # amsgrad=True variant with gradients that shrink, so that the running maximum differs from the moving average
    q = [torch.nn.Parameter(out["p0:0"].clone()), torch.nn.Parameter(out["p0:1"].clone())]
    opt = ref.novograd.Novograd(q, lr=0.01, betas=(0.95, 0.5), weight_decay=1e-3, amsgrad=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for step in range(3):
            for i, t in enumerate(q):
                t.grad = out["g%d:%d" % (step, i)].clone() * (0.5 ** step)
            opt.step()
            for i, t in enumerate(q):
                out["ams_p%d:%d" % (step + 1, i)] = t.detach().clone()
    np.savez_compressed(os.path.join(OUT, "novograd.npz"), **_np(out))


def gen_features(ref):
    """SpectrogramExtractor.extract of the reference on seeded synthetic audio (two lengths); the dither noise it drew is
    reproduced from the same torch seed and stored, so that consumers can feed it explicitly."""
    from oracle.ref_loader import load_reference_features, to_attr
    dl = load_reference_features()
    conf = to_attr(dict(sample_rate=16000, window_size=0.02, window_stride=0.01, window="hamming"))
    ex = dl.SpectrogramExtractor(conf, mel_spec=64)
    out = {"fb": ex.fb[0].numpy()}
    rs = np.random.RandomState(7)
    for name, n in (("a", 16000), ("b", 11237)):
        t = np.arange(n) / 16000.0
        sig = (0.3 * np.sin(2 * np.pi * (200 + 900 * t) * t) + 0.05 * rs.randn(n)).astype(np.float32)
        torch.manual_seed(11 + n)
        feats = ex.extract(sig)
        torch.manual_seed(11 + n)
        noise = torch.randn(sig.shape)
        out[name + ":signal"], out[name + ":noise"], out[name + ":feats"] = sig, noise.numpy(), feats.numpy()
    np.savez_compressed(os.path.join(OUT, "features.npz"), **out)


def gen_beam(ref):
    """prefix_beam_search of the reference (decoder.py:147-231) on seeded random posteriors: string and score."""
    labels = list(ref.label_sets.labels_map["english_lowercase"])
    rs = np.random.RandomState(21)
    out = {"labels": np.array(labels)}
    cases = [("sharp", 40, 4.0, 5, 5, 1e-3, False), ("flat", 25, 0.5, 20, 5, 1e-3, False), ("k1", 30, 2.0, 1, 5, 1e-3, False),
             ("beta0", 30, 2.0, 5, 0, 0.0, False), ("betaf", 30, 2.0, 5, 2.5, 0.05, False), ("f32", 35, 3.0, 5, 5, 1e-3, True),
             ("lm", 30, 3.0, 5, 5, 1e-3, False)]
    for name, T, temp, k, beta, prune, f32 in cases:
        x = rs.randn(T, len(labels)) * temp
        x[:, 28] += 1.0
        p = np.exp(x) / np.exp(x).sum(1, keepdims=True)
        if f32:
            p = p.astype(np.float32)
        lm = (lambda s: 1.0 / (1.0 + len(s))) if name == "lm" else None
        string, score = ref.decoder.prefix_beam_search(p, labels, 0, lm, k, 0.3, beta, prune, return_weights=True)
        out[name + ":probs"], out[name + ":params"] = p, np.array([k, beta, prune], dtype=np.float64)
        out[name + ":string"], out[name + ":score"] = np.array(string), np.float64(score)
    np.savez_compressed(os.path.join(OUT, "beam.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] in ("features", "beam", "novograd", "strided", "narrow", "odd", "jasper_odd"):      # regenerate one fixture without touching the others
        ref = rl.load_reference()
        torch.set_num_threads(1)
        {"features": gen_features, "beam": gen_beam, "novograd": gen_novograd, "strided": gen_strided, "narrow": gen_narrow, "odd": gen_odd, "jasper_odd": gen_jasper_odd}[sys.argv[1]](ref)
        return
    ref = rl.load_reference()
    torch.set_num_threads(1)
    gen_decoder(ref)
    gen_conv_block(ref)
    gen_w2l(ref)
    gen_jasper(ref)
    gen_jasper_dense(ref)
    gen_strided(ref)
    gen_narrow(ref)
    gen_odd(ref)
    gen_jasper_odd(ref)
    gen_ctc(ref)
    gen_novograd(ref)
    gen_features(ref)
    gen_beam(ref)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
