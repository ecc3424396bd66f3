import sys, os, torch
sys.path.insert(0, os.getcwd())
from miphei_vit_b200 import ops
M, D = 10528, 1536
xn = torch.randn(M, D + 64, device="cuda").bfloat16(); dq = torch.randn(M, 3 * D + 64, device="cuda").bfloat16()
o = [torch.zeros(D, 8, device="cuda"), torch.zeros(D, 8, device="cuda"), torch.zeros(8, D, device="cuda"), torch.zeros(8, D, device="cuda")]
ws = torch.empty(int(ops._lib.load().mv_lora_grads_workspace_bytes(M, D)), dtype=torch.uint8, device="cuda")
f = lambda: ops.lora_grads(xn, dq, D, 0.5, o[0], o[1], o[2], o[3], workspace=ws)
for _ in range(3): f()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(20): f()
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): g.replay()
e1.record(); torch.cuda.synchronize()
print("lora_grads M=%d D=%d: %.1f us" % (M, D, e0.elapsed_time(e1) / 100 * 1e3))
