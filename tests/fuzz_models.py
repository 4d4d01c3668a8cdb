"""TEST INFRASTRUCTURE, run by hand where /root/reference is mounted: whole-model differential fuzzing against the UNMODIFIED reference.

    python tests/fuzz_models.py w2l|jasper [--seed N] [--cases K]

For every case a random model configuration (widths, kernel sizes, strides, dilations; Jasper: repeats, residuals, separable blocks,
masks) and a random ragged batch are drawn; the reference itself (oracle/ref_loader.py) produces the fixture in memory exactly as
oracle/gen_golden.py freezes the committed ones -- train-mode forward, CTC loss, every parameter gradient, running statistics, eval
forward, greedy transcripts -- and the GPU model tests' own checkers (tests/test_gpu_models.py check_w2l_golden / check_jasper_golden,
same tolerances) run this package against it on the emulated GPU (tests/_fake_cuda.py: unmodified host package, the library's C
wrappers and kernels compiled for the host).  The five committed fixtures are points of this space; this walks the rest of it."""
import argparse
import json
import os
import random
import sys
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import gen_golden as G  # noqa: E402
from oracle import ref_loader as rl  # noqa: E402


def _randomise_bn(model):
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.2, 0.2)


def _batch(rnd, B, T, S, mask_inputs):
    x = torch.randn(B, 64, T)
    il = torch.tensor([T] + [rnd.randint(max(T // 2, 40), T) for _ in range(B - 1)], dtype=torch.int32)
    tl = torch.tensor([S] + [rnd.randint(1, S) for _ in range(B - 1)], dtype=torch.int32)
    tg = torch.randint(1, 29, (B, S), dtype=torch.int32)
    for n in range(B):
        tg[n, tl[n]:] = 0
        if mask_inputs:
            x[n, :, il[n]:] = 0
    return x, il, tg, tl


def w2l_case(ref, rnd):
    n = rnd.choice([2, 3, 3, 4])
    layers = []
    for i in range(n):
        k = rnd.choice([1, 3, 5, 7, 11, 13])
        layers.append(dict(output_size=rnd.choice([64, 64, 72, 80, 128, 136]), kernel_size=k, stride=2 if (i == 0 or rnd.random() < 0.2) else 1,
                           dilation=rnd.choice([1, 1, 1, 2]) if k > 1 else 1, dropout=-1))
    cfg = rl.reference_model_cfg("wav2letter", mid_layers=n, dropout=-1)
    cfg["layers"] = rl.to_attr(layers)
    model = ref.wav2letter.Wav2Letter(cfg)
    _randomise_bn(model)
    x, il, tg, tl = _batch(rnd, rnd.choice([2, 3]), rnd.choice([161, 201, 240]), rnd.choice([6, 12]), False)
    out = {"x": x, "il": il, "tg": tg, "tl": tl,
           "layers": np.array([[l["output_size"], l["kernel_size"], l["stride"], l["dilation"]] for l in layers])}
    out.update(G._sd(model, "sd0:"))
    model.train()
    out.update({"train:" + k: v for k, v in G._train_step_record(ref, model, x, il, tg, tl, None).items()})
    out.update(G._sd(model, "sd1:"))
    model.eval()
    with torch.no_grad():
        o, ol = model(x, il)
    out["eval:out"], out["eval:out_len"] = o, ol
    out["eval:decoded"] = np.array(model.ctc_decoder.decode(o, ol))
    out["scaling_factor"] = model.scaling_factor
    return layers, G._np(out)


def jasper_case(ref, rnd, seed):
    n = rnd.choice([2, 3, 4])
    blocks = []
    for i in range(n):
        stride = 2 if (i == 0 or rnd.random() < 0.15) else 1
        sep = rnd.random() < 0.3
        k = rnd.choice([3, 5, 6, 10, 11]) if not sep else rnd.choice([5, 11, 12, 33])
        b = dict(layer_size=rnd.choice([64, 64, 72, 80, 128, 136]), kernel_size=k, stride=stride, residual=(stride == 1 and i > 0 and rnd.random() < 0.6),
                 separable=sep, repeat=rnd.choice([1, 1, 2, 3]))
        if stride == 1 and not sep and rnd.random() < 0.2:
            b["dilation"] = 2
        blocks.append(b)
    torch.manual_seed(seed)
    cfg = rl.reference_model_cfg("jasper", mid_layers=n, dropout=0, jasper_blocks=blocks)
    model = ref.jasper.Jasper(cfg)
    _randomise_bn(model)
    x, il, tg, tl = _batch(rnd, rnd.choice([2, 3]), rnd.choice([201, 281, 401]), rnd.choice([4, 8]), True)
    out = {"x": x, "il": il, "tg": tg, "tl": tl, "blocks_json": np.array(json.dumps(blocks))}
    out.update(G._sd(model, "sd0:"))
    model.train()
    out.update({"train:" + k: v for k, v in G._train_step_record(ref, model, x, il, tg, tl, None).items()})
    out.update({k: v for k, v in G._sd(model, "sd1:").items() if "running" in k or "num_batches" in k})
    model.eval()
    with torch.no_grad():
        o, ol = model(x, il)
    out["eval:out"], out["eval:out_len"] = o, ol
    out["scaling_factor"] = model.scaling_factor
    return blocks, G._np(out)


class _Npz(dict):
    @property
    def files(self):
        return list(self.keys())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("family", choices=["w2l", "jasper"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cases", type=int, default=10)
    a = ap.parse_args()
    if not rl.reference_available():
        raise SystemExit("the reference tree is not mounted")
    ref = rl.load_reference()
    fixtures = []
    rnd = random.Random(a.seed)
    for c in range(a.cases):                              # fixtures first: the reference runs on plain torch, before "cuda" is faked
        torch.manual_seed(1000 * a.seed + c)
        try:
            fixtures.append(w2l_case(ref, rnd) if a.family == "w2l" else jasper_case(ref, rnd, 1000 * a.seed + c))
        except Exception as e:  # noqa: BLE001  (a configuration the reference itself rejects, e.g. a kernel wider than the signal)
            print("reference rejected a configuration:", repr(e)[:200])
    import _fake_cuda
    _fake_cuda.enable()
    import wav2letter_pytorch_b200 as pkg
    import test_gpu_models as T
    failures = 0
    for c, (conf, g) in enumerate(fixtures):
        if float(g["train:loss"]) == 0.0:                 # every utterance infeasible (total stride too large for the targets): the
            print("skip", conf, "(reference loss is exactly 0: the checker's relative tolerance is undefined)", flush=True)   # checkers compare relatively
            continue
        try:
            if a.family == "w2l":
                # bf16 storage error grows with depth (the committed 3-block fixture holds 3e-2 against the bf16-emulating oracle); the
                # checker's second criterion -- no worse against the fp32 reference than 1.5x the emulation's own error -- stays as is
                T.check_w2l_golden(pkg, _Npz(g), emu_tol=3e-2 * 2 ** max(0, len(conf) - 2))
            else:
                T.check_jasper_golden(pkg, _Npz(g), seed=1000 * a.seed + c, emu_tol=0.25)
            print("ok  ", conf, flush=True)
        except Exception:  # noqa: BLE001
            failures += 1
            print("FAIL", conf, flush=True)
            traceback.print_exc(limit=3)
    print("done, failures:", failures, "of", len(fixtures))
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
