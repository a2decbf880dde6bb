"""(python scripts/perf_aux.py [--only loss|head|ref])  GPU timings of the other hot-path rows (loss fwd/bwd at cfg3, projection head) and of the informal comparator
(the reference's own einsum/max/sum chain executed by PyTorch on the same B200).  Prints one JSON line each."""
import json, sys, time
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
import colpali_b200 as cb
from oracle import li_oracle as O

dev = torch.device("cuda:0")
PEAK_HBM = 6569.6
try:
    PEAK_HBM = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass

def cuda_time(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

# ---- loss at cfg3 ------------------------------------------------------------------------------
q, d, lens = O.cfg3_inputs()
q, d = q.to(dev), d.to(dev)
for name, mod in (("ColbertLoss", cb.ColbertLoss()), ("ColbertPairwiseCELoss", cb.ColbertPairwiseCELoss())):
    fwd = cuda_time(lambda: mod(q, d))
    qq, dd = q.clone().requires_grad_(True), d.clone().requires_grad_(True)
    def step():
        qq.grad = None; dd.grad = None
        mod(qq, dd).backward()
    fb = cuda_time(step)
    print(json.dumps({"what": f"{name} cfg3 (B=64, Nq=32, Nd<=1030 left-padded)", "fwd_ms": fwd, "fwd_bwd_ms": fb,
                      "fwd_tflops": 2 * 64 * 32 * 64 * 1030 * 128 / fwd / 1e9}), flush=True)

# reference chain on the same GPU (informal comparator)
def ref_loss():
    return O.colbert_loss_port(q, d)
t_ref = cuda_time(ref_loss, n=5, warm=2)
qq, dd = q.clone().requires_grad_(True), d.clone().requires_grad_(True)
def ref_step():
    qq.grad = None; dd.grad = None
    O.colbert_loss_port(qq, dd).backward()
t_ref_fb = cuda_time(ref_step, n=5, warm=2)
print(json.dumps({"what": "reference ColbertLoss chain (torch einsum/amax/CE) on the same B200, bf16", "fwd_ms": t_ref, "fwd_bwd_ms": t_ref_fb}), flush=True)

# ---- scorer comparator: reference score_multi_vector chain on the GPU ------------------------------
qs, ps = O.cfg2_inputs()
qs, ps = qs.to(dev), ps.to(dev)
def ref_score():
    return O.score_multi_vector_port(qs, ps, batch_size=128, device=dev)
t = cuda_time(ref_score, n=3, warm=1)
print(json.dumps({"what": "reference score_multi_vector chain (torch einsum/max/sum, batch 128) on the same B200, device-resident inputs",
                  "ms": t, "queries_per_s": 32 / t * 1e3}), flush=True)

# ---- head -----------------------------------------------------------------------------------------
for tokens, hidden in ((275 * 1000, 1536), (1030 * 64, 2048)):
    h = torch.randn(tokens, hidden, device=dev).bfloat16()
    lin = torch.nn.Linear(hidden, 128).to(dev, torch.bfloat16)
    mask = torch.ones(tokens, dtype=torch.long, device=dev)
    t = cuda_time(lambda: cb.fused_head(h, lin.weight, lin.bias, mask))
    bytes_alg = 2 * tokens * (hidden + 128) + 2 * hidden * 128
    t_ref = cuda_time(lambda: O.head_port(h, lin.weight, lin.bias, mask), n=5, warm=2)
    print(json.dumps({"what": f"fused_head T={tokens} H={hidden}", "ms": t, "gbs": bytes_alg / t / 1e6,
                      "frac_hbm": bytes_alg / t / 1e6 / PEAK_HBM, "tokens_per_s": tokens / t * 1e3,
                      "reference_chain_same_gpu_ms": t_ref}), flush=True)
