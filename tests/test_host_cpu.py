"""CPU-only tests: the C-ABI library loads and exports every symbol the header declares, argument validation works
without a GPU, the host-side mirror of the reference interface behaves like the reference (config composition, label
tables, padding rule, parameter layout, state_dict contract, error behaviour), and the data-parallel reducer is
exercised with world_size=2 over gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from wav2letter_pytorch_b200 import _lib
    _lib.build()
    return _lib.load()


def test_abi_exports_every_declared_symbol(lib):
    from wav2letter_pytorch_b200 import _lib
    header = open(os.path.join(ROOT, "include", "w2l_sm100.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(w2l_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), "missing export: %s" % name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.w2l_version() >= 100
    assert lib.w2l_launch_count() == 0                      # nothing computes on the CPU box


def test_abi_argument_validation_without_gpu(lib):
    from wav2letter_pytorch_b200 import _lib
    from wav2letter_pytorch_b200._lib import ConvDesc
    d = ConvDesc(2, 100, 60, 64, 64, 3, 1, 100, 0, 100, 0, 64, 0, 0)       # Cin = 60 < 64
    buf = ctypes.c_void_p(0x1000)
    rc = lib.w2l_conv1d_fwd(buf, buf, None, None, None, None, buf, ctypes.byref(d), None)
    assert rc == 1 and b"Cin" in lib.w2l_last_error()
    assert lib.w2l_conv1d_fwd(buf, buf, None, None, None, None, buf, None, None) == 1
    assert lib.w2l_greedy_decode(None, 2, 10, 29, 290, 29, None, 0, None, None, None, None, None, 0, None) == 1
    assert lib.w2l_ctc_loss(buf, 0, 2, 10, 500, 5000, 500, buf, 4, buf, buf, 0, 1, 1, None, None, None, buf, 1 << 20, None) == 1
    assert lib.w2l_ctc_loss_workspace_bytes(64, 750, 225) > 64 * 750 * 451 * 4
    assert lib.w2l_ctc_loss_workspace_bytes(1, 10, 100000) == 0          # lattice too long -> unsupported
    assert lib.w2l_greedy_decode_workspace_bytes(64, 750) == 64 * 3 * 4
    with pytest.raises(_lib.W2LError):
        _lib.check(1, "probe")
    # host Levenshtein
    a = np.array([1, 2, 3, 4], dtype=np.int32)
    b = np.array([1, 3, 4, 5, 6], dtype=np.int32)
    P = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    assert lib.w2l_edit_distance_host(P(a), 4, P(b), 5) == 3
    assert lib.w2l_edit_distance_host(P(a), 0, P(b), 5) == 5


def test_no_cpu_fallback():
    from wav2letter_pytorch_b200 import config, functional as F
    from wav2letter_pytorch_b200.ctc_loss import CTCLoss
    from wav2letter_pytorch_b200.wav2letter import Wav2Letter
    with pytest.raises(RuntimeError):
        F.greedy_decode(torch.rand(1, 4, 5))
    with pytest.raises(RuntimeError):
        CTCLoss(zero_infinity=True)(torch.randn(5, 1, 4).log_softmax(-1), torch.ones(1, 2, dtype=torch.int32), torch.tensor([5]), torch.tensor([2]))
    model = Wav2Letter(config.compose().model)
    with pytest.raises(RuntimeError):
        model(torch.randn(1, 64, 50), torch.tensor([50]))
    src = "".join(open(os.path.join(ROOT, "wav2letter_pytorch_b200", f)).read() for f in os.listdir(os.path.join(ROOT, "wav2letter_pytorch_b200"))
                  if f.endswith(".py"))
    assert "oracle" not in src.replace("oracle/", "")          # the product never imports the checker


def test_config_compose_and_labels():
    from wav2letter_pytorch_b200 import config, label_sets
    cfg = config.compose()
    assert cfg.model.name == "wav2letter" and cfg.model.mid_layers == 1 and len(cfg.model.layers) == 20
    assert cfg.model.input_size == 64 and cfg.data.mel_spec == 64 and cfg.data.audio_conf == cfg.model.audio_conf
    assert cfg.model.labels == label_sets.english_lowercase_labels and len(cfg.model.labels) == 29
    assert cfg.model.labels[0] == "_" and cfg.model.labels[28] == " " and cfg.model.labels[1] == "'"
    assert cfg.model.decoder["_target_"] == "decoder.GreedyDecoder" and cfg.model.decoder.labels == cfg.model.labels
    assert cfg.model.optimizer["_target_"] == "torch.optim.SGD" and cfg.model.optimizer.nesterov is True
    j = config.compose(overrides=["model=jasper", "model.mid_layers=15", "trainer.gpus=8"])
    assert j.model.name == "jasper" and len(j.model.jasper_blocks) == 15 and j.trainer.gpus == 8
    assert [b.kernel_size for b in j.model.jasper_blocks][:5] == [32, 32, 32, 32, 38]
    j10 = config.compose(overrides=["model=jasper10x5", "optimizer=novograd"])
    assert sum(b.get("repeat", 1) for b in j10.model.jasper_blocks) == 53 and j10.model.optimizer["_target_"] == "novograd.Novograd"
    assert len(label_sets.hebrew_labels) == 29 and len(label_sets.english_labels) == 29
    dec = config.instantiate(cfg.model.decoder)
    assert type(dec).__name__ == "GreedyDecoder" and dec.space_index == 28 and dec.blank_index == 0


from wav2letter_pytorch_b200.layers import ConvParams as _CP
ConvParamsPhys = _CP.phys


def test_padding_rule_and_model_contract():
    from oracle import w2l_oracle as O
    from wav2letter_pytorch_b200 import config
    from wav2letter_pytorch_b200.wav2letter import Conv1dBlock, Wav2Letter, reflect_padding
    # SURVEY 8a-1 table
    assert reflect_padding(64, 11, 2, 1) == (4, 5) and reflect_padding(256, 11, 1, 1) == (5, 5)
    assert reflect_padding(768, 29, 1, 2) == (28, 28) and reflect_padding(896, 1, 1, 1) == (0, 0)
    for cin in (33, 64, 161):
        for k in (1, 4, 11):
            for s in (1, 2, 3):
                for d in (1, 2):
                    assert reflect_padding(cin, k, s, d) == O.reflect_pad_amounts(cin, k, s, d)
    model = Wav2Letter(config.compose(overrides=["model.mid_layers=20"]).model)
    assert sum(p.numel() for p in model.parameters()) == 153074845              # SURVEY section 0.1 [probe]
    assert model.scaling_factor == 2
    il = torch.tensor([1001, 800], dtype=torch.int32)
    out_len = model.compute_output_lengths(il)
    assert out_len.tolist() == [500, 400] and out_len.dtype == torch.int32
    keys = list(model.state_dict().keys())
    assert keys[0] == "conv1ds.conv1d_0.conv1.weight" and "conv1ds.conv1d_20.conv1.bias" in keys
    assert "conv1ds.conv1d_19.batch_norm.num_batches_tracked" in keys and "conv1ds.conv1d_20.batch_norm.weight" not in keys
    blk = model.conv1ds.conv1d_16
    assert blk.conv1.weight.shape == (896, 768, 29) and blk.pad_lr == (28, 28) and blk.batch_norm.momentum == 0.9
    assert model.conv1ds.conv1d_15.next_pad == (28, 28) and model.conv1ds.conv1d_19.next_pad == (0, 0)
    assert Wav2Letter(config.compose().model).conv1ds.conv1d_1.conv1.out_channels == 29         # literal default: 1 block + head
    odd = Conv1dBlock(161, 250, (11,), 2)                     # any width, as in the reference: padded internally (ConvParams.phys)
    assert odd.conv1.weight.shape == (250, 161, 11) and odd.conv1.padded
    assert (odd.conv1.cin_phys, odd.conv1.cout_phys, odd.conv1.cout_pad) == (11 * 168, 256, 256)
    assert not Conv1dBlock(64, 256, (11,), 2).conv1.padded and not model.conv1ds.conv1d_16.conv1.padded
    assert ConvParamsPhys(36) == 64 and ConvParamsPhys(250) == 256 and ConvParamsPhys(896) == 896


def test_conv_params_layout_roundtrip():
    from wav2letter_pytorch_b200.layers import ConvParams
    torch.manual_seed(0)
    ref = torch.nn.Conv1d(8, 16, 5)
    torch.manual_seed(0)
    for unfold in (False, True):
        torch.manual_seed(0)
        c = ConvParams(8, 16, 5, stride=2 if unfold else 1, unfold=unfold)
        assert torch.equal(c.weight.detach(), ref.weight.detach()) and torch.equal(c.bias.detach(), ref.bias.detach())   # same init stream
        st = c.storage()
        assert st.is_contiguous() and st.data_ptr() == c.weight.data_ptr()
        if unfold:
            assert st.shape == (1, 16, 40) and torch.equal(st[0].view(16, 5, 8).permute(0, 2, 1), c.weight.detach())
        else:
            assert st.shape == (5, 16, 8) and torch.equal(st.permute(1, 2, 0), c.weight.detach())
        dw = torch.randn_like(st)
        gv = c.grad_view(dw)
        assert gv.shape == c.weight.shape and gv.stride() == c.weight.stride()
        w2 = torch.randn(16, 8, 5)
        c.load_state_dict({"weight": w2, "bias": torch.zeros(16)})
        assert torch.equal(c.weight.detach(), w2) and c.storage().data_ptr() == c.weight.data_ptr()
        import copy
        c2 = copy.deepcopy(c)
        assert c2.weight.stride() == c.weight.stride() and c2.storage().is_contiguous()
    assert ConvParams(64, 29, 1).cout_pad == 64 and ConvParams(64, 896, 3).cout_pad == 896


def test_decoder_host_logic():
    from wav2letter_pytorch_b200.decoder import Decoder, GreedyDecoder, _edit_distance
    d = Decoder(["_", "a", "b", " "])
    assert d.int_to_char == {0: "_", 1: "a", 2: "b", 3: " "} and d.space_index == 3
    assert Decoder(["_", "a"]).space_index == 2                      # no space -> out-of-range index, as the reference
    assert d.cer("ab ba", "abba") == 0 and d.wer("ab ba", "ab b a") == 2
    assert d.cer_ratio("a b", "b") == (1, 2) and d.wer_ratio("a b c", "a c") == (1, 3)
    assert _edit_distance("kitten", "sitting") == 3 and _edit_distance("", "") == 0
    with pytest.raises(NotImplementedError):
        d.decode(None)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            GreedyDecoder(["_", "a", "b", " "]).decode(torch.rand(1, 3, 4))


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
from wav2letter_pytorch_b200.distributed import GradientReducer, broadcast_buffers, shard_batch, dense_view
from wav2letter_pytorch_b200.layers import ConvParams, BatchNormParams
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
torch.manual_seed(0)
class Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = ConvParams(4, 6, 3)
        self.bn = BatchNormParams(6)
net = Net()
red = GradientReducer(net)
# a stand-in loss that touches every parameter through its reference-shaped view
def loss_of(scale):
    return scale * ((net.conv.weight ** 2).sum() + net.conv.bias.sum() + (net.bn.weight * net.bn.bias).sum() + net.bn.weight.sum())
loss_of(float(rank + 1)).backward()
red.finish()
expect = (1.0 + 2.0) / 2
assert torch.allclose(net.conv.weight.grad, 2 * expect * net.conv.weight.detach(), atol=1e-6), rank
assert net.conv.weight.grad.stride() == net.conv.weight.stride()
assert torch.allclose(net.conv.bias.grad, torch.full((6,), expect))
assert dense_view(net.conv.weight.grad).is_contiguous()
net.bn.running_mean.fill_(float(rank + 5))
broadcast_buffers(net)
assert float(net.bn.running_mean[0]) == 5.0
b = shard_batch((torch.arange(8).view(8, 1), torch.arange(8), None), rank, world)
assert b[1].tolist() == list(range(rank * 4, rank * 4 + 4)) and b[2] is None
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_gradient_reducer_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER % {"root": ROOT})
    port = 29500 + os.getpid() % 2000
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "rank %d ok" % r in o, o


def test_mel_filterbank_matches_oracle_restatement():
    """two independent restatements of librosa.filters.mel (vectorised in the package, scalar loops in the oracle)"""
    from oracle import w2l_oracle as O
    from wav2letter_pytorch_b200.features import mel_filterbank
    for sr, n_fft, n_mels in ((16000, 512, 64), (8000, 256, 40), (22050, 2048, 128)):
        a, b = mel_filterbank(sr, n_fft, n_mels, 0.0, sr / 2), O.mel_filterbank_slaney(sr, n_fft, n_mels, 0.0, sr / 2)
        assert a.shape == (n_mels, n_fft // 2 + 1) and a.dtype == np.float32
        assert np.allclose(a, b, rtol=0, atol=1e-7)
        assert (a >= 0).all() and (a.sum(1) > 0).all()          # every filter has support


def test_spectrogram_extractor_needs_cuda():
    import torch
    from wav2letter_pytorch_b200.features import SpectrogramExtractor
    ex = SpectrogramExtractor(dict(sample_rate=16000, window_size=0.02, window_stride=0.01, window="hamming"), mel_spec=64)
    assert ex.n_fft == 512 and ex.win_length == 320 and ex.hop_length == 160 and ex.fb.shape == (1, 64, 257)
    assert ex.n_frames(16000) == 101
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            ex.extract(np.zeros(16000, dtype=np.float32))


# ---------------------------------------------------------------------------------------------- prefix beam search (host routine)
@pytest.mark.parametrize("name", ["sharp", "flat", "k1", "beta0", "betaf", "f32", "lm"])
def test_beam_search_library_golden(golden, name):
    """library routine vs the reference's own outputs: identical transcript AND identical float64 score"""
    from wav2letter_pytorch_b200.decoder import prefix_beam_search
    g = golden("beam")
    labels = [str(s) for s in g["labels"]]
    k, beta, prune = g[name + ":params"]
    beta = int(beta) if float(beta).is_integer() else float(beta)
    lm = (lambda s: 1.0 / (1.0 + len(s))) if name == "lm" else None
    string, score = prefix_beam_search(g[name + ":probs"], labels, 0, lm, int(k), 0.3, beta, float(prune), return_weights=True)
    assert string == str(g[name + ":string"])
    assert score == float(g[name + ":score"])


def test_beam_search_library_vs_oracle_random():
    from oracle import w2l_oracle as O
    from wav2letter_pytorch_b200.decoder import PrefixBeamSearchLMDecoder, prefix_beam_search
    rs = np.random.RandomState(5)
    labels = ["_", "'", "a", "b", "c", " ", ">"]                      # includes the end character: prefixes that end are frozen
    batch = []
    for trial in range(25):
        T = int(rs.randint(2, 40))
        x = rs.randn(T, len(labels)) * rs.choice([0.7, 2.0, 4.0])
        p = np.exp(x) / np.exp(x).sum(1, keepdims=True)
        k, beta, prune = int(rs.choice([1, 2, 5, 12])), rs.choice([5, 1, 0.5]), float(rs.choice([1e-3, 0.1, 0.0]))
        beta = int(beta) if float(beta).is_integer() else float(beta)
        lm = (lambda s: 0.5 + 0.1 * (len(s) % 3)) if trial % 5 == 0 else None
        want = O.prefix_beam_search(p, labels, 0, lm, k, 0.3, beta, prune)
        got = prefix_beam_search(p, labels, 0, lm, k, 0.3, beta, prune, return_weights=True)
        assert got[0] == want[0] and got[1] == want[1], (trial, got, want)
        if T == 17 or len(batch) < 3:
            batch.append(p)
    # batch through the Decoder interface (threads inside the library): same strings as utterance by utterance
    T = min(b.shape[0] for b in batch)
    probs = np.stack([b[:T] for b in batch])
    dec = PrefixBeamSearchLMDecoder(None, labels, blank_index=0, k=4, alpha=0.3, beta=5, prune=1e-3)
    assert dec.decode(probs) == [O.prefix_beam_search(b, labels, 0, None, 4, 0.3, 5, 1e-3)[0] for b in probs]
    assert dec.decode(torch.from_numpy(probs[0]).float()) == O.prefix_beam_search(probs[0].astype(np.float32), labels, 0, None, 4, 0.3, 5, 1e-3)[0]
    with pytest.raises(NotImplementedError):
        dec.decode(probs, return_offsets=True)
    with pytest.raises(AssertionError):
        prefix_beam_search(-probs[0], labels)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the contract's keys; under a
    multi-rank launch only rank 0 prints.  Tiny sample here (mid_layers=1, one utterance) so the CPU suite stays fast."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--mid-layers", "1",
           "--cpu-batch", "1"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "audio-sec/sec per train step" and line["unit"] == "audio-s/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] == 1
    assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=60, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_bench_gpu_arm_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=120)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


def test_get_time_per_word_matches_reference_semantics():
    from wav2letter_pytorch_b200.decoder import get_time_per_word
    pred = "hi  there a"
    offs = [3, 5, 9, 10, 12, 13, 15, 18, 20, 22, 30]
    assert get_time_per_word(pred, offs, ratio=0.02) == [("hi", 3 * 0.02, 5 * 0.02), ("there", 12 * 0.02, 20 * 0.02), ("a", 30 * 0.02, 30 * 0.02)]
    assert get_time_per_word("", []) == [] and get_time_per_word("  ", [1, 2]) == []
    try:                                                       # the reference function itself, where its tree is mounted
        from oracle import ref_loader
        if ref_loader.reference_available():
            ref = ref_loader.load_reference()
            for p, o in ((pred, offs), ("a b", [0, 1, 2]), (" x", [4, 7]), ("ab", [1, 2])):
                assert get_time_per_word(p, o, 0.5) == ref.decoder.get_time_per_word(p, o, 0.5)
    except ImportError:
        pass


def test_jasper_unmasked_block_speaks_reference_state_dict_keys():
    """conv_mask=False: the reference holds a bare nn.Conv1d (jasper.py:289-298), key `mconv.N.weight`; masked blocks say
    `mconv.N.conv.weight` (jasper.py:96-105).  Checkpoints must load and save under those names."""
    from wav2letter_pytorch_b200.jasper import JasperBlock
    blk = JasperBlock(64, 64, repeat=2, kernel_size=5, residual=True, separable=False, conv_mask=False, activation=torch.nn.ReLU())
    keys = set(blk.state_dict().keys())
    assert {"mconv.0.weight", "mconv.4.weight", "res.0.0.weight"} <= keys and not any(".conv.weight" in k for k in keys)
    sd = {k: torch.randn_like(v) if v.is_floating_point() else v for k, v in blk.state_dict().items()}
    missing, unexpected = blk.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    assert torch.equal(blk.mconv[0].conv.weight.detach(), sd["mconv.0.weight"])
    masked = JasperBlock(64, 64, repeat=1, kernel_size=5, residual=False, separable=True, conv_mask=True, activation=torch.nn.ReLU())
    assert {"mconv.0.conv.weight", "mconv.1.conv.weight"} <= set(masked.state_dict().keys())
    with pytest.raises(NotImplementedError):                 # activation-then-dropout differs from the fused order for a clamp
        JasperBlock(64, 64, repeat=1, kernel_size=5, dropout=0.2)


def test_encode_refs_into_caller_buffers():
    """decoder._encode_refs(into=...): what graph_step.GraphedTrainStep uses to refresh the static reference operands of a captured
    step -- label ids, lengths and the three denominators of base_asr_models.py:58-69 (characters without spaces, words, characters)"""
    from wav2letter_pytorch_b200.decoder import GreedyDecoder
    from wav2letter_pytorch_b200.label_sets import english_lowercase_labels as labels
    dec = GreedyDecoder(labels)
    texts = ["hello world", " a  b ", "it's", "x"]
    ids = torch.full((4, 16), -7, dtype=torch.int32)
    lens = torch.zeros(4, dtype=torch.int32)
    enc = dec._encode_refs(texts, into=(ids, lens))
    assert enc is not None
    slot, smax, cer_den, wer_den, len_den = enc
    assert slot[0] is ids and slot[1] is lens and smax == 11
    assert lens.tolist() == [len(t) for t in texts]
    for i, t in enumerate(texts):
        assert ids[i, :len(t)].tolist() == [labels.index(c) for c in t]
        assert (ids[i, len(t):] == -7).all()                                  # nothing beyond the transcript is touched
    assert cer_den == sum(len(t.replace(" ", "")) for t in texts)
    assert wer_den == sum(len(t.split()) for t in texts)
    assert len_den == sum(len(t) for t in texts)
    assert dec._encode_refs(texts, into=(ids[:, :8], lens)) is None          # a transcript longer than the buffer: no device path
    assert dec._encode_refs(texts + ["y"], into=(ids, lens)) is None         # more transcripts than rows
