"""BASELINE configs[3]: corpus-sharded scoring, 128 queries x (docs_per_gpu x world) documents of 1030 tokens.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/run_cfg4.py [--docs-per-gpu 12500]

Every rank generates its shard on the device (seed 1000 + rank, never materialised on the host), plants each query's
positive document in a known shard, scores with the fused kernel and merges the per-rank top-10 with one NCCL all-gather.
Checks, all against something that is NOT the kernel:
  (a) the planted positives must be the merged top-1 of every query (asserted);
  (b) every rank's local [128, docs_per_gpu] score matrix against fp32 torch (matmul / amax / sum on the same GPU) --
      max relative error, row argmax and local top-10 ids (the reference's arithmetic, processing_utils.py:179, in fp32);
  (c) the merged top-10 against the merged top-10 of those fp32 matrices (recall@10 asserted, >= 0.998: see below).
Prints one JSON line on rank 0.
"""
import argparse, json, os, sys
import torch
import torch.distributed as dist
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import colpali_b200 as cb
from colpali_b200.sharded import merge_topk, score_sharded

N_Q, N_TOK, N_D, DIM, K = 128, 32, 1030, 128, 10


def make_queries(dev):
    g = torch.Generator(device=dev).manual_seed(0)
    return F.normalize(torch.randn(N_Q, N_TOK, DIM, device=dev, generator=g), dim=-1).bfloat16()


def make_shard(rank, world, docs_per_gpu, q, dev):
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    bank = torch.empty(docs_per_gpu, N_D, DIM, dtype=torch.bfloat16, device=dev)
    for lo in range(0, docs_per_gpu, 500):
        hi = min(lo + 500, docs_per_gpu)
        bank[lo:hi] = F.normalize(torch.randn(hi - lo, N_D, DIM, device=dev, generator=g), dim=-1).bfloat16()
    gn = torch.Generator(device=dev).manual_seed(7)
    noise = torch.randn(N_Q, N_TOK, DIM, device=dev, generator=gn)
    planted = {}
    for i in range(N_Q):
        if i % world == rank:
            j = (i * 97) % docs_per_gpu
            bank[j, -N_TOK:] = F.normalize(q[i].float() + 0.08 * noise[i], dim=-1).bfloat16()  # cos ~ 0.74 to its query token
            planted[i] = rank * docs_per_gpu + j
    return bank, planted


def fp32_scores(q, docs, chunk=50):
    """Plain PyTorch fp32 MaxSim on the GPU (no TF32), chunked over documents."""
    torch.backends.cuda.matmul.allow_tf32 = False
    nq, nt, dim = q.shape
    q2 = q.float().reshape(nq * nt, dim)
    out = torch.empty(nq, docs.shape[0], dtype=torch.float32, device=q.device)
    for lo in range(0, docs.shape[0], chunk):
        d = docs[lo:lo + chunk].float()
        s = q2 @ d.reshape(-1, dim).t()
        out[:, lo:lo + chunk] = s.view(nq, nt, d.shape[0], d.shape[1]).amax(3).sum(1)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs-per-gpu", type=int, default=12500)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    q = make_queries(dev)
    shard, planted = make_shard(rank, world, a.docs_per_gpu, q, dev)
    bank = cb.DocBank.from_passages(shard, dev)
    total = a.docs_per_gpu * world
    lo = rank * a.docs_per_gpu
    for _ in range(2):
        ts, ti = score_sharded(q, bank, lo, total, top_k=K)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        ts, ti = score_sharded(q, bank, lo, total, top_k=K)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
    # kernel alone (local shard, no top-k / exchange)
    qb = cb.QueryBlock(q, dev)
    e0.record()
    for _ in range(a.steps):
        local_scores = cb.maxsim(qb, bank)
    e1.record(); torch.cuda.synchronize()
    ms_kernel = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX); dist.all_reduce(ms_kernel, op=dist.ReduceOp.MAX)

    # (b) this rank's whole score matrix against fp32 torch
    want = fp32_scores(q, shard)
    rel = ((local_scores - want).abs() / want.abs().clamp_min(1e-3)).max()
    local_ok = torch.tensor([float(rel), float(torch.equal(local_scores.argmax(1), want.argmax(1))),
                             float(torch.equal(local_scores.topk(K, 1).indices, want.topk(K, 1).indices))], device=dev)
    # (c) merged top-k of the fp32 matrices, through the same exchange
    ws, wi = score_sharded(q, bank, lo, total, top_k=K, local_scorer=lambda _q, _b: want)
    all_planted = [None] * world
    stats = [torch.zeros_like(local_ok) for _ in range(world)]
    if world > 1:
        dist.all_gather_object(all_planted, planted)
        dist.all_gather(stats, local_ok)
    else:
        all_planted, stats = [planted], [local_ok]
    if rank == 0:
        truth = {k: v for d in all_planted for k, v in d.items()}
        ti_c, wi_c = ti.cpu(), wi.cpu()
        top1 = sum(int(ti_c[i, 0]) == truth[i] for i in range(N_Q)) / N_Q
        rec = sum(truth[i] in ti_c[i].tolist() for i in range(N_Q)) / N_Q
        recall_fp32 = float(sum(len(set(ti_c[i].tolist()) & set(wi_c[i].tolist())) for i in range(N_Q)) / (N_Q * K))
        st = torch.stack(stats).cpu()
        out = {"config": "cfg4 corpus-sharded scoring", "world": world, "docs_per_gpu": a.docs_per_gpu, "queries": N_Q,
               "ms_per_batch": float(ms), "queries_per_s": N_Q / float(ms) * 1e3,
               "kernel_ms": float(ms_kernel),
               "tflops_per_gpu_kernel": 2.0 * N_Q * N_TOK * a.docs_per_gpu * N_D * DIM / float(ms_kernel) / 1e9,
               "planted_top1": top1, "planted_recall_at_10": rec,
               "recall_at_10_vs_fp32_torch": recall_fp32, "top10_ids_equal_fp32_torch": bool(torch.equal(ti_c, wi_c)),
               "per_rank_max_rel_err_vs_fp32_torch": [float(x) for x in st[:, 0]],
               "per_rank_argmax_equal_fp32_torch": [bool(x) for x in st[:, 1]],
               "per_rank_top10_equal_fp32_torch": [bool(x) for x in st[:, 2]]}
        print(json.dumps(out), flush=True)
        assert top1 == 1.0 and rec == 1.0, "a planted positive is not the merged top-1"
        # (two random documents whose fp32 scores differ by less than the kernel's ~4e-7 relative error may swap places
        # at the top-10 boundary: allow two such swaps in 1280 ids, report exact equality separately)
        assert recall_fp32 >= 0.998, "merged top-10 differs from the fp32 torch scorer's"
        assert float(st[:, 0].max()) < 1e-4, "a local score differs from fp32 torch by more than 1e-4 relative"
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
