"""One 8-GPU-sized dense shard (2,626,916 x 768 bf16, B=4096, k=100) on one GPU: where does a call's time go?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsearch_b200 import _native as nat  # noqa: E402
from vsearch_b200.index import _Engine  # noqa: E402

dev = torch.device("cuda:0")
n, d, B, k = 2_626_916, 768, 4096, 100
g = torch.Generator(device=dev).manual_seed(7)
x = torch.randn((n, d), generator=g, device=dev).to(torch.bfloat16)
q = torch.randn((B, d), generator=g, device=dev).to(torch.bfloat16)
eng = _Engine.from_dense(x, dev, torch.bfloat16)
for _ in range(2):
    eng.search(q, k, score_round=nat.VS_BF16)
torch.cuda.synchronize()
eng.kernel_timer(reset=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    eng.search(q, k, score_round=nat.VS_BF16)
e1.record()
torch.cuda.synchronize()
ms, nl = eng.kernel_timer(reset=True)
print("call ms", e0.elapsed_time(e1) / 3, "sweeps+merges ms", ms / 3, "TFLOPs(call)", 2.0 * B * n * d / (e0.elapsed_time(e1) / 3 * 1e-3) / 1e12)
