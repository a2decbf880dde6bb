"""(python scripts/one_shot.py dense|topk)  One launch of a late-round kernel, for `ncu --set full` captures."""
import sys
import torch
sys.path.insert(0, ".")
import colpali_b200 as cb
from colpali_b200.scoring import DocBank, QueryBlock, maxsim_topk

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
unit = lambda *s: torch.nn.functional.normalize(torch.randn(*s, generator=g, device=dev), dim=-1)
if sys.argv[1] == "dense":
    a, b = unit(1000, 1536), unit(100000, 1536)
    for _ in range(2):
        cb.score_single_vector(a, b, device=dev)
else:
    q, bank = QueryBlock(unit(128, 32, 128).bfloat16(), dev), DocBank.from_passages(unit(12500, 1030, 128).bfloat16(), dev)
    for _ in range(2):
        maxsim_topk(q, bank, 10)
torch.cuda.synchronize()
