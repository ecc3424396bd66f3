"""torchrun --nproc-per-node N tools/dp_check.py : data-parallel consistency on real GPUs (NCCL)."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from miphei_vit_b200.generators.mipheivit import get_vitmatte
from miphei_vit_b200.trainer import Trainer
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.manual_seed(0)
with torch.device("cuda"):
    m = get_vitmatte("hoptimus0", 256, 16, use_lora=True, embed_dim=256, depth=3, num_heads=4, hidden=512)
m = m.cuda()
with torch.no_grad():
    for blk in m.encoder.vit.blocks:
        blk.ls1.gamma.fill_(0.3); blk.ls2.gamma.fill_(0.3)
        blk.attn.qkv.lora_q.B.normal_(0, 0.02); blk.attn.qkv.lora_v.B.normal_(0, 0.02)
# weights must start identical on every rank
tr = Trainer(m, marker_weights=torch.linspace(1, 10, 16), batch_size=4, total_steps=100, warmup_steps=2)
ref = tr.flat.clone(); dist.broadcast(ref, 0); assert torch.equal(ref, tr.flat), "initial weights differ"
g = torch.Generator(device="cpu").manual_seed(77 + rank)
x = torch.randn((4, 3, 256, 256), generator=g).cuda()
y = (torch.rand((4, 16, 256, 256), generator=g) * 1.8 - 0.9).cuda()
losses = []
for _ in range(3):
    losses.append(tr.step(x, y).item())
torch.cuda.synchronize()
allp = [torch.empty_like(tr.flat) for _ in range(world)]
dist.all_gather(allp, tr.flat)
same = all(torch.equal(allp[0], a) for a in allp)
allg = [torch.empty_like(tr.gflat) for _ in range(world)]
dist.all_gather(allg, tr.gflat)
sameg = all(torch.equal(allg[0], a) for a in allg)
print("rank %d losses %s params identical across ranks: %s grads identical: %s moved: %.3e" % (
    rank, ["%.4f" % l for l in losses], same, sameg, float((tr.flat - ref).abs().max())), flush=True)
assert same and sameg and all(l == l for l in losses)
dist.destroy_process_group()
