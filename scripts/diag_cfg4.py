"""Diagnosis of the round-1 cfg4 planted-recall failure: kernel scores vs an fp32 torch scorer at the cfg4 geometry.

    python scripts/diag_cfg4.py [--docs 12500] [--queries 128] [--out gpurun_out/diag_cfg4.json]

One GPU.  Builds the cfg4 shard exactly as scripts/run_cfg4.py does (rank 0 of a world of `--world`), scores it with the
fused kernel under several launch geometries (balanced on/off, cluster 1/2, query tiles per CTA 1/2) and compares the
FULL [queries, docs] matrix with torch fp32 (matmul in chunks, amax, sum).  Prints where the mismatches are.
"""
import argparse, json, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import colpali_b200 as cb
from colpali_b200 import _lib
from scripts.run_cfg4 import make_queries, make_shard, N_D, N_TOK, DIM


def torch_scores(q, bank, chunk=50):
    """fp32 reference on the GPU: [n_q, n_docs] = sum_n max_s <q, d> (processing_utils.py:179 in fp32)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    nq, nt, _ = q.shape
    q2 = q.float().reshape(nq * nt, DIM)
    out = torch.empty(nq, bank.shape[0], dtype=torch.float32, device=q.device)
    for lo in range(0, bank.shape[0], chunk):
        d = bank[lo:lo + chunk].float()                      # [c, L, D]
        s = q2 @ d.reshape(-1, DIM).t()                       # [nq*nt, c*L]
        out[:, lo:lo + chunk] = s.view(nq, nt, d.shape[0], d.shape[1]).amax(3).sum(1)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=12500)
    ap.add_argument("--world", type=int, default=2)
    ap.add_argument("--out", default="gpurun_out/diag_cfg4.json")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    q = make_queries(dev)
    shard, planted = make_shard(0, a.world, a.docs, q, dev)
    want = torch_scores(q, shard)
    res = {"docs": a.docs, "queries": int(q.shape[0]), "planted": len(planted), "variants": {}}
    pl_q = torch.tensor(sorted(planted), device=dev)
    pl_d = torch.tensor([planted[i] for i in sorted(planted)], device=dev)
    res["torch_planted_top1"] = float((want.argmax(1)[pl_q] == pl_d).float().mean())
    res["torch_planted_score_mean"] = float(want[pl_q, pl_d].mean())
    res["torch_score_median"] = float(want.median())
    bank = cb.DocBank.from_passages(shard, dev)
    qb = cb.QueryBlock(q, dev)
    for name, opts in [("default", {}), ("unbalanced", {"balanced": 0}), ("cluster1", {"cluster": 1}),
                       ("r1", {"qtiles_per_cta": 1}), ("r1_cluster1_unbalanced", {"qtiles_per_cta": 1, "cluster": 1, "balanced": 0})]:
        for k, v in opts.items():
            _lib.set_option(k, v)
        got = cb.maxsim(qb, bank)
        torch.cuda.synchronize()
        for k in opts:
            _lib.set_option(k, 1 if k == "balanced" else 0)
        err = (got - want).abs() / want.abs().clamp_min(1e-3)
        bad = err > 1e-4
        v = {"max_rel_err": float(err.max()), "n_bad": int(bad.sum()), "n_total": bad.numel(),
             "argmax_equal": bool(torch.equal(got.argmax(1), want.argmax(1))),
             "planted_top1": float((got.argmax(1)[pl_q] == pl_d).float().mean()),
             "planted_score_mean": float(got[pl_q, pl_d].mean()),
             "nan": int(torch.isnan(got).sum()), "inf": int(torch.isinf(got).sum())}
        if bad.any():
            bq, bd = bad.nonzero(as_tuple=True)
            v["bad_queries"] = sorted(set(bq.tolist()))[:40]
            v["bad_docs_first"] = sorted(set(bd.tolist()))[:40]
            v["bad_docs_count"] = len(set(bd.tolist()))
            v["bad_per_query_hist"] = torch.bincount(bq, minlength=q.shape[0]).tolist()
            k = min(10, bq.numel())
            v["examples"] = [(int(bq[i]), int(bd[i]), float(got[bq[i], bd[i]]), float(want[bq[i], bd[i]])) for i in range(k)]
        res["variants"][name] = v
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res)[:4000])


if __name__ == "__main__":
    main()
