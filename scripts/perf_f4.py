"""(python scripts/perf_f4.py)  GPU timings of the late-round additions: projection head after the 64-token-unit schedule,
the top-k selection fused into the MaxSim kernel's tail, the bi-encoder losses / single-vector scorer / similarity maps on
the dense fp32 kernel -- each next to what it replaces, executed by PyTorch on the same B200.  One JSON line each."""
import json, sys
import torch
sys.path.insert(0, ".")
import colpali_b200 as cb
from colpali_b200.scoring import DocBank, QueryBlock, maxsim, maxsim_topk
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


def graph_time(fn, n=20, reps=5):
    """GPU time per call from CUDA-graph replays of n calls (no host launch path between the kernels)."""
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (n * reps)


def interleaved(fns, rounds=12, per=4):
    """Median over rounds of each variant's time, the variants taking turns inside every round (the SM clock drifts by
    several percent within seconds: blocks measured one after the other are not comparable)."""
    for f in fns:
        for _ in range(2): f()
    torch.cuda.synchronize()
    ts = [[] for _ in fns]
    for _ in range(rounds):
        for i, f in enumerate(fns):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(per): f()
            e1.record(); torch.cuda.synchronize()
            ts[i].append(e0.elapsed_time(e1) / per)
    return [sorted(t)[len(t) // 2] for t in ts]


which = set(sys.argv[1:]) or {"head", "topk", "bi"}
if "head" in which:
    for tokens, hidden in ((275 * 1000, 1536), (1030 * 64, 2048), (1030 * 64, 1536), (34125, 1536)):
        h = torch.randn(tokens, hidden, device=dev).bfloat16()
        lin = torch.nn.Linear(hidden, 128).to(dev, torch.bfloat16)
        mask = torch.ones(tokens, dtype=torch.long, device=dev)
        t_host = cuda_time(lambda: cb.fused_head(h, lin.weight, lin.bias, mask), n=50, warm=5)
        t = graph_time(lambda: cb.fused_head(h, lin.weight, lin.bias, mask))
        bytes_alg = 2 * tokens * (hidden + 128) + 2 * hidden * 128
        print(json.dumps({"what": f"fused_head T={tokens} H={hidden} (64-token unit shares)", "gpu_ms_graph": t,
                          "ms_with_python_launch_path": t_host, "gbs": bytes_alg / t / 1e6,
                          "frac_hbm": bytes_alg / t / 1e6 / PEAK_HBM, "tokens_per_s": tokens / t * 1e3}), flush=True)
        del h

if "topk" in which:
    g = torch.Generator(device=dev).manual_seed(0)
    unit = lambda *s: torch.nn.functional.normalize(torch.randn(*s, generator=g, device=dev), dim=-1).bfloat16()
    for n_q, n_docs in ((128, 12500), (32, 1000)):
        qs, ps = unit(n_q, 32, 128), unit(n_docs, 1030, 128)
        q, bank = QueryBlock(qs, dev), DocBank.from_passages(ps, dev)
        t0, t1, t2 = interleaved([lambda: maxsim(q, bank), lambda: maxsim_topk(q, bank, 10),
                                  lambda: torch.topk(maxsim(q, bank), 10, dim=1)])
        print(json.dumps({"what": f"local top-10 of {n_q} q x {n_docs} docs x 1030 x 128", "scores_only_ms": t0,
                          "fused_topk_ms": t1, "scores_then_torch_topk_ms": t2, "tail_us": (t1 - t0) * 1e3,
                          "torch_topk_us": (t2 - t0) * 1e3}), flush=True)
        del ps, bank

if "bi" in which:
    g = torch.Generator(device=dev).manual_seed(1)
    unit = lambda *s: torch.nn.functional.normalize(torch.randn(*s, generator=g, device=dev), dim=-1)
    for b, c, dim, dt in ((64, 512, 1536, torch.bfloat16), (64, 64, 1536, torch.float32)):
        q, d = unit(b, dim).to(dt), unit(c, dim).to(dt)
        for name, mod, ref in (("BiEncoderLoss", cb.BiEncoderLoss(), lambda x, y: O.bi_loss_port("ce", x, y)),):
            fwd = cuda_time(lambda: mod(q, d), n=50)
            qq, dd = q.clone().requires_grad_(True), d.clone().requires_grad_(True)
            def step():
                qq.grad = None; dd.grad = None
                mod(qq, dd).backward()
            fb = cuda_time(step, n=50)
            rq, rd = q.clone().requires_grad_(True), d.clone().requires_grad_(True)
            def ref_step():
                rq.grad = None; rd.grad = None
                torch.nn.functional.cross_entropy(torch.einsum("bd,cd->bc", rq, rd) / 0.02, torch.arange(b, device=dev)).backward()
            rfb = cuda_time(ref_step, n=50)
            print(json.dumps({"what": f"{name} B={b} C={c} D={dim} {str(dt)[6:]}", "fwd_ms": fwd, "fwd_bwd_ms": fb,
                              "reference_chain_same_gpu_fwd_bwd_ms": rfb}), flush=True)
    a, bmat = unit(1000, 1536), unit(100000, 1536)
    t = cuda_time(lambda: cb.score_single_vector(a, bmat, device=dev), n=5, warm=1)
    tr = cuda_time(lambda: torch.einsum("bd,cd->bc", a, bmat), n=5, warm=1)
    print(json.dumps({"what": "score_single_vector 1000 x 100000 x 1536 fp32", "ms": t, "tflops_fp32": 2 * 1000 * 100000 * 1536 / t / 1e9,
                      "torch_einsum_same_gpu_ms": tr}), flush=True)
    img, qe = unit(8, 1030, 128).bfloat16(), unit(8, 20, 128).bfloat16()
    mask = torch.zeros(8, 1030, dtype=torch.bool, device=dev); mask[:, 4:1028] = True
    t = cuda_time(lambda: cb.get_similarity_maps_from_embeddings(img, qe, (32, 32), mask), n=10)
    print(json.dumps({"what": "similarity maps, 8 pages x 20 query tokens x 32 x 32 patches", "ms": t}), flush=True)
