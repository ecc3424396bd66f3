"""Isolates the decoder: hand-written train fwd/bwd vs torch fp32 autograd on identical inputs and weights."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import model as om
from miphei_vit_b200 import ops
from miphei_vit_b200.generators.mipheivit import get_vitmatte
from miphei_vit_b200.autograd import decoder_forward_torch
from miphei_vit_b200.decoder_train import DecoderTrain
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
S, D, C, B = 128, 128, 16, 2
cfg = om.Config(img_size=S, embed_dim=D, depth=1, num_heads=2, hidden=256, out_chans=C)
sd = om.init_state_dict(cfg, seed=5, perturb=True)
m = get_vitmatte("hoptimus0", S, C, use_lora=True, embed_dim=D, depth=1, num_heads=2, hidden=256)
m.load_state_dict(sd); m = m.cuda().train()
eng = m.engine
eng._ensure_packed(train=True)
x = om.normalize_tiles(om.synthetic_tiles_u8(B, S, seed=1)).cuda()
g = torch.Generator(device="cuda").manual_seed(3)
fmap = torch.randn((B, S // 16, S // 16, D), generator=g, device="cuda").bfloat16()
dpred = torch.randn((B, C, S, S), generator=g, device="cuda") * 1e-3
# torch reference on the same (bf16-valued) inputs
img = x.bfloat16().float()
feat = fmap.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
import copy
dec_ref = copy.deepcopy(m.decoder)
pr = decoder_forward_torch(dec_ref, feat, img)
pr.backward(dpred)
ref = {n: p.grad for n, p in dec_ref.named_parameters()}
# hand-written
ws = eng._workspace(B)
ws.x_in.copy_(x)
ops.prep_input(ws.x_in, img=ws.img8, want_patches=False)
dt = DecoderTrain(eng); dt.pack()
pred = dt.forward(fmap, ws.img8)
grads, dfmap = dt.backward(dpred)
print("pred pearson %.7f maxabs %.2e" % (om.pearson(pred.cpu(), pr.detach().cpu()), (pred - pr.detach()).abs().max().item()))
print("dfeat cosine %.6f" % om.cosine(dfmap.float().permute(0, 3, 1, 2).cpu(), feat.grad.cpu()))
names = {p: n for n, p in m.decoder.named_parameters()}
for p, gval in grads.items():
    n = names[p]
    if "segmentation_head" in n and "head_0." not in n: continue
    r = ref[n]
    print("%-50s cos %.6f |g| %.3e ref %.3e" % (n, om.cosine(gval.float().cpu(), r.cpu()), float(gval.norm()), float(r.norm())))
