import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsearch_b200.index import _Engine
from vsearch_b200 import _native as nat
dev = torch.device("cuda:0")
n, d, B, k = 6_000_000, 768, 2048, 100
g = torch.Generator(device=dev).manual_seed(7)
x = torch.randn((n, d), generator=g, device=dev).to(torch.bfloat16)
q = torch.randn((B, d), generator=g, device=dev).to(torch.bfloat16)
eng = _Engine.from_dense(x, dev, torch.bfloat16)
for _ in range(2):
    ids, sc = eng.search(q, k, score_round=nat.VS_BF16)
torch.cuda.synchronize()
ms, nl = eng.kernel_timer(reset=True)
print("kernel ms", ms / 2, "TFLOPs", 2.0 * B * n * d / (ms / 2 * 1e-3) / 1e12)
