import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import w2l_oracle as O
from wav2letter_pytorch_b200 import config
from wav2letter_pytorch_b200.jasper import Jasper
g = np.load("tests/golden/jasper_dense.npz")
blocks = [dict(b, dropout=0) for b in json.loads(str(g["blocks_json"]))]
cfg = config.compose(overrides=["model=jasper", "model.mid_layers=5"]).model
cfg["jasper_blocks"] = config.to_attr(blocks)
model = Jasper(cfg)
model.load_state_dict({k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd0:")})
model.cuda().train()
x, il, tg, tl = (torch.from_numpy(g[k]).cuda() for k in ("x", "il", "tg", "tl"))
out, ol = model(x, il)
loss = model.criterion(out.transpose(0, 1), tg, ol, tl)
loss.backward()
specs = O.jasper_block_specs(blocks)
sd = {k[4:]: torch.from_numpy(g[k]).clone() for k in g.files if k.startswith("sd0:")}
ep = {k: v.requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
e_out, e_ol = O.jasper_forward(x.cpu(), il.cpu(), sd, specs, True, emu=True)
e_loss = torch.nn.CTCLoss(blank=0, reduction="mean", zero_infinity=True)(e_out.transpose(0, 1), tg.cpu(), e_ol, tl.cpu())
e_loss.backward()
rl = lambda a, b: float((a.double().cpu() - b.double().cpu()).norm() / (b.double().cpu().norm() + 1e-30))
print("out vs emu", rl(out.detach(), e_out.detach()), "loss", loss.item(), e_loss.item(), float(g["train:loss"]))
for n, p in model.named_parameters():
    ref = torch.from_numpy(g["train:grad:" + n])
    print("%-45s cuda-emu %.4f  cuda-ref %.4f  emu-ref %.4f" % (n, rl(p.grad, ep[n].grad), rl(p.grad, ref), rl(ep[n].grad, ref)))
