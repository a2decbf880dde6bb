"""GPU parity: the fused sm_100a MaxSim kernel, called through the C ABI, against the CPU oracle, the
reference-generated golden vectors, and size-independent properties at the full BASELINE size."""
import pytest
import torch

from conftest import from_bits, load_golden, split_rows

import colpali_b200 as cb
from oracle import li_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
REL_TOL = 1e-2  # north_star: MaxSim score within 1e-2 rel-tol (observed ~1e-6: fp32 accumulate of exact bf16 products)


def rel_err(got, want):
    return ((got - want).abs() / want.abs().clamp_min(1e-3)).max().item()


def test_reference_unit_test_shapes_list_equals_tensor():
    g = load_golden("scorer_small.npz")
    q_pad, p_pad = torch.from_numpy(g["t1_q"]), torch.from_numpy(g["t1_p"])
    qs = [q_pad[i, : int(n)] for i, n in enumerate(g["t1_qlen"])]
    ps = [p_pad[j, : int(n)] for j, n in enumerate(g["t1_plen"])]
    s_list = cb.score_multi_vector(qs, ps, device=DEV)
    s_tensor = cb.score_multi_vector(q_pad, p_pad, device=DEV)
    assert s_list.shape == (2, 3) and s_list.dtype == torch.float32 and s_list.device.type == "cpu"
    assert torch.allclose(s_list, s_tensor)  # tests/utils/test_processing_utils.py:26-35
    # the kernel contracts bf16: compare tightly with the oracle on bf16-rounded inputs, loosely with fp32 golden
    want = O.score_multi_vector_port([x.bfloat16().float() for x in qs], [x.bfloat16().float() for x in ps])
    assert torch.allclose(s_list, want, rtol=1e-5, atol=1e-4)
    assert torch.allclose(s_list, torch.from_numpy(g["t1_list"]), rtol=2e-2, atol=0.15)


def test_zero_padding_semantics():
    g = load_golden("scorer_small.npz")
    q = [torch.tensor([[1.0, 0.0]])]
    a = torch.tensor([[-1.0, 0.0], [-0.5, 0.0]])
    b3 = torch.tensor([[-0.2, 0.0], [-0.3, 0.0], [-0.9, 0.0]])
    # -0.2 / -0.3 / -0.9 are not bf16-representable: the kernel contracts bf16, hence the 2e-3 slack
    close = lambda x, y: torch.allclose(x, y, rtol=0, atol=2e-3)  # noqa: E731
    assert close(cb.score_multi_vector(q, [a], device=DEV), torch.from_numpy(g["t2_alone"]))            # -0.5
    assert close(cb.score_multi_vector(q, [a, b3], device=DEV), torch.from_numpy(g["t2_batched"]))       # 0.0, -0.2
    assert cb.score_multi_vector(q, [a, b3], device=DEV)[0, 0] == 0.0  # shorter doc: zero pad row wins the max
    assert close(cb.score_multi_vector(q, [a, b3], batch_size=1, device=DEV), torch.from_numpy(g["t2_bs1"]))


@pytest.mark.parametrize("batch_size,key", [(128, "t3_fp32"), (3, "t3_fp32_bs3")])
def test_ragged_against_reference_golden(batch_size, key):
    g = load_golden("scorer_small.npz")
    qs = split_rows(from_bits(g["t3_q"], (-1, 128)), g["t3_qlen"])  # N_q in {5,32,17,1,40}: nq_pad = 64
    ps = split_rows(from_bits(g["t3_p"], (-1, 128)), g["t3_plen"])  # N_d in {1..1030}
    got = cb.score_multi_vector(qs, ps, batch_size=batch_size, device=DEV)
    want = torch.from_numpy(g[key])
    assert rel_err(got, want) < 1e-4
    assert torch.equal(got.argmax(1), want.argmax(1))


def test_ragged_bf16_rounding_mode_matches_native_bf16_reference():
    g = load_golden("scorer_small.npz")
    qs = split_rows(from_bits(g["t3_q"], (-1, 128)), g["t3_qlen"])
    ps = split_rows(from_bits(g["t3_p"], (-1, 128)), g["t3_plen"])
    got = cb.score_multi_vector(qs, ps, device=DEV, round_bf16=True)
    want = torch.from_numpy(g["t3_bf16"])
    # identical up to one bf16 ulp (the reference rounds every dot product, we round the per-token maximum)
    ulp = want.abs().clamp_min(1e-3) * 2.0 ** -7
    assert ((got - want).abs() <= ulp).all()
    assert (got == want).float().mean() > 0.9


def test_cfg1_scores_and_argmax():
    g = load_golden("scorer_cfg1.npz")
    q, d = O.cfg1_inputs()
    got = cb.score_multi_vector(q, d, device=DEV)
    want = torch.from_numpy(g["ref_fp32"])
    assert rel_err(got, want) < 1e-5
    assert torch.equal(got.argmax(1), want.argmax(1))
    assert rel_err(got, torch.from_numpy(g["ref_bf16"])) < REL_TOL


def test_cfg2_headline_scores_argmax_and_ties():
    """32 queries x 1000 docs x 1030 x 128 bf16 (BASELINE configs[1]) against the reference's outputs."""
    g = load_golden("scorer_cfg2.npz")
    q, d = O.cfg2_inputs()
    got = cb.score_multi_vector(q, d, device=DEV)
    ref32 = torch.from_numpy(g["ref_fp32"])
    ref16 = torch.from_numpy(g["ref_bf16"])
    assert got.shape == (32, 1000)
    assert rel_err(got, ref32) < 1e-5 < REL_TOL
    assert torch.equal(got.argmax(1), ref32.argmax(1))  # bit-exact argmax-doc agreement (fp32 reference)
    assert rel_err(got, ref16) < REL_TOL
    # native-bf16 reference rounds scores -> ties; our argmax must sit inside its tied-max set
    am = got.argmax(1)
    assert (ref16[torch.arange(32), am] == ref16.max(1).values).all()
    # rounding mode reproduces the bf16 reference to within one bf16 ulp
    got16 = cb.score_multi_vector(q, d, device=DEV, round_bf16=True)
    assert ((got16 - ref16).abs() <= ref16.abs() * 2.0 ** -7).all()
    assert (got16 == ref16).float().mean() > 0.9


def test_full_size_properties():
    """Size-independent checks at the full cfg2 size, no oracle needed."""
    q, d = O.cfg2_inputs()
    dev = torch.device(DEV)
    qd, dd = q.to(dev), d.to(dev)
    bank = cb.DocBank.from_passages(dd, dev)
    base = cb.maxsim(cb.QueryBlock(qd, dev), bank)
    # (1) permuting documents permutes columns, bit-exactly
    perm = torch.randperm(1000, generator=torch.Generator().manual_seed(3)).to(dev)
    s_perm = cb.maxsim(cb.QueryBlock(qd, dev), cb.DocBank.from_passages(dd[perm], dev))
    assert torch.equal(s_perm, base[:, perm])
    # (2) scaling queries by 2 scales scores by exactly 2 (power-of-two scaling is exact in bf16/fp32)
    s2 = cb.maxsim(cb.QueryBlock(qd * 2, dev), bank)
    assert torch.equal(s2, base * 2)
    # (3) splitting a query's tokens splits its score additively
    sa = cb.maxsim(cb.QueryBlock(qd[:, :16], dev), bank)
    sb = cb.maxsim(cb.QueryBlock(qd[:, 16:], dev), bank)
    assert torch.allclose(sa + sb, base, rtol=1e-6, atol=1e-5)
    # (4) a document split in two: per-token max of the halves combines by max -> score(whole) <= sum, >= each
    h1 = cb.maxsim(cb.QueryBlock(qd, dev), cb.DocBank.from_passages(dd[:, :515], dev))
    h2 = cb.maxsim(cb.QueryBlock(qd, dev), cb.DocBank.from_passages(dd[:, 515:], dev))
    assert (base >= torch.maximum(h1, h2) - 1e-5).all() and (base <= h1 + h2 + 1e-5).all()
    # (5) query order does not matter; different query-tile grouping (R) gives identical bits
    s_one = torch.cat([cb.maxsim(cb.QueryBlock(qd[i : i + 4], dev), bank) for i in range(0, 32, 4)])
    assert torch.equal(s_one, base)
    # (6) argmax variant returns the same scores and indices that reproduce them
    s_arg, am = cb.maxsim(cb.QueryBlock(qd, dev), bank, want_argmax=True)
    assert torch.equal(s_arg, base)
    assert am.shape == (1000, 32 * 32) and int(am.min()) >= 0 and int(am.max()) < 1030
    j = 123
    sim = qd.float().reshape(-1, 128) @ dd[j].float().T  # [1024, 1030]
    assert torch.equal(am[j].long(), sim.argmax(1))
    assert torch.allclose(sim.gather(1, am[j].long()[:, None]).view(32, 32).sum(1), base[:, j], rtol=1e-6, atol=1e-5)


def test_argmax_on_ragged_with_floor():
    g = load_golden("scorer_small.npz")
    dev = torch.device(DEV)
    qs = [x.to(dev) for x in split_rows(from_bits(g["t3_q"], (-1, 128)), g["t3_qlen"])]
    ps = [x.to(dev) for x in split_rows(from_bits(g["t3_p"], (-1, 128)), g["t3_plen"])]
    bank = cb.DocBank.from_passages(ps, dev)
    qb = cb.QueryBlock(qs, dev)
    s, am = cb.maxsim(qb, bank, want_argmax=True)
    assert torch.equal(s, cb.maxsim(qb, bank))
    for j, p in enumerate(ps):
        for i, q in enumerate(qs):
            sim = q.float() @ p.float().T
            mx, ix = sim.max(1)
            rows = am[j, i * qb.nq_pad : i * qb.nq_pad + q.shape[0]].long()
            floor_hit = rows < 0
            assert torch.equal(floor_hit, mx <= 0) or bank.floor is None
            assert torch.equal(rows[~floor_hit], ix[~floor_hit])


def test_many_query_tiles_and_wide_queries():
    dev = torch.device(DEV)
    g = torch.Generator().manual_seed(9)
    qs = [torch.nn.functional.normalize(torch.randn(n, 128, generator=g), dim=-1).bfloat16() for n in [70] * 7 + [96, 33]]
    ps = [torch.nn.functional.normalize(torch.randn(n, 128, generator=g), dim=-1).bfloat16() for n in [300, 511, 513, 64, 1]]
    got = cb.score_multi_vector(qs, ps, device=dev)
    want = torch.from_numpy(O.maxsim_f64(qs, ps, O.reference_floors([300, 511, 513, 64, 1]))).float()
    assert rel_err(got, want) < 1e-4


def test_single_vector_scorer():
    g = torch.Generator().manual_seed(2)
    qs = [torch.randn(32, generator=g) for _ in range(4)]
    ps = [torch.randn(32, generator=g) for _ in range(8)]
    got = cb.score_single_vector(qs, ps, device=DEV)
    want = torch.einsum("bd,cd->bc", torch.stack(qs), torch.stack(ps))  # fp32 operands stay fp32 (tests/test_bi_gpu.py)
    assert got.shape == (4, 8) and got.dtype == torch.float32
    assert torch.allclose(got.cpu(), want, rtol=1e-5, atol=1e-5)


def test_errors():
    with pytest.raises(ValueError, match="No queries"):
        cb.score_multi_vector([], [torch.randn(3, 128)], device=DEV)
    with pytest.raises(ValueError, match="No passages"):
        cb.score_multi_vector([torch.randn(3, 128)], [], device=DEV)
    with pytest.raises(cb.ColpaliB200Error):
        cb.score_multi_vector([torch.randn(3, 400)], [torch.randn(3, 400)], device=DEV)


def _torch_fp32_maxsim(q, ps, floors, dev):
    """Plain PyTorch fp32 evaluation on the GPU (padded einsum + masked amax): fast checker for the shape sweep."""
    lens = torch.tensor([p.shape[0] for p in ps], device=dev)
    pad = torch.nn.utils.rnn.pad_sequence([p.to(dev).float() for p in ps], batch_first=True)  # [n_d, Lmax, 128]
    if pad.shape[1] == 0:
        pad = torch.zeros(len(ps), 1, 128, device=dev)
    sim = torch.einsum("bnd,csd->bcns", q.to(dev).float(), pad)
    valid = torch.arange(pad.shape[1], device=dev)[None, :] < lens[:, None]
    sim = sim.masked_fill(~valid[None, :, None, :], float("-inf"))
    mx = sim.amax(dim=3)
    if floors is not None:
        mx = torch.maximum(mx, torch.tensor(floors, device=dev)[None, :, None])
    return mx.sum(dim=2).cpu()


@pytest.mark.parametrize("n_queries", [3, 40])
def test_tile_balanced_partitions_on_a_ragged_bank(n_queries):
    """Many short ragged documents (some empty): partitions are whole tiles, documents are cut at partition
    boundaries and re-assembled through the split workspace; the first document of a partition is found by binary
    search.  Scores (with the reference's padding floors) and argmax must match the fp64 evaluation."""
    dev = torch.device(DEV)
    g = torch.Generator().manual_seed(12)
    lens = torch.randint(0, 300, (700,), generator=g).tolist()
    lens[0] = 0
    lens[17] = 0
    lens[699] = 0
    qs = [O.unit_rows((32, 128), 500 + i) for i in range(n_queries)]
    ps = [O.unit_rows((n, 128), 1000 + j) if n else torch.zeros(0, 128, dtype=torch.bfloat16) for j, n in enumerate(lens)]
    bank = cb.DocBank.from_passages([p.to(dev) for p in ps], dev)
    assert bank.contiguous and bank.max_len == max(lens) and bank.uniform_len == 0
    qb = cb.QueryBlock([x.to(dev) for x in qs], dev)
    got, am = cb.maxsim(qb, bank, want_argmax=True)
    got2 = cb.maxsim(qb, bank)
    floors = O.reference_floors(lens, 128)
    want = _torch_fp32_maxsim(torch.stack(qs), ps, floors, dev)
    finite = torch.isfinite(want)
    assert torch.equal(torch.isfinite(got.cpu()), finite)
    assert torch.allclose(got.cpu()[finite], want[finite], rtol=1e-5, atol=1e-4)
    assert torch.equal(got, got2)
    # argmax of a few documents against a direct evaluation
    for j in (1, 100, 350, 698):
        if lens[j] == 0:
            continue
        sim = torch.cat(qs).float() @ ps[j].float().T
        mx, ix = sim.max(1)
        rows = am[j].cpu().long()
        hit_floor = rows < 0
        assert torch.equal(rows[~hit_floor], ix[~hit_floor])
        assert (mx[hit_floor] <= 0).all()
    # switching balancing off gives bit-identical scores
    from colpali_b200 import _lib
    _lib.set_option("balanced", 0)
    try:
        assert torch.equal(cb.maxsim(qb, bank), got2)
    finally:
        _lib.set_option("balanced", 1)


@pytest.mark.parametrize("n_queries", [12, 20, 36])
def test_odd_number_of_query_tile_groups(n_queries):
    """3, 5 and 9 query tiles: the last 2-CTA cluster of a partition has a CTA without query tiles, which must still
    keep the shared TMA ring moving."""
    dev = torch.device(DEV)
    qs = O.unit_rows((n_queries, 32, 128), 77)
    ps = O.unit_rows((90, 300, 128), 78)
    got = cb.score_multi_vector(qs, ps, device=dev)
    want = torch.einsum("bnd,csd->bcns", qs.float(), ps.float()).amax(3).sum(2)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("seed", range(8))
def test_randomised_shapes_against_fp32_torch(seed):
    """Seeded sweep over query counts / lengths, bank sizes, ragged and dense banks, empty documents: exercises the
    R = 1 / 2 variants, odd group counts, clusters with a phantom CTA, balanced and whole-document partitions."""
    dev = torch.device(DEV)
    g = torch.Generator().manual_seed(1000 + seed)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))  # noqa: E731
    n_q, nq = ri(1, 50), ri(1, 70)
    n_d = ri(1, 500)
    dense = seed % 3 == 0
    if dense:
        lens = [ri(1, 400)] * n_d
    else:
        lens = [ri(0, 350) if ri(0, 9) else 0 for _ in range(n_d)]
        if sum(lens) == 0:
            lens[0] = 5
    q = O.unit_rows((n_q, nq, 128), 3 * seed)
    bank = O.unit_rows((sum(lens), 128), 7000 + seed)
    ps = list(torch.split(bank, lens))
    if dense:
        got = cb.score_multi_vector(q, torch.stack(ps), device=dev)
        floors = None
    else:
        got = cb.score_multi_vector(list(q), ps, device=dev)
        floors = O.reference_floors(lens, 128)
    want = _torch_fp32_maxsim(q, ps, floors, dev)
    finite = torch.isfinite(want)
    assert torch.equal(torch.isfinite(got), finite), (n_q, nq, n_d, dense)
    assert torch.allclose(got[finite], want[finite], rtol=1e-5, atol=2e-4), (n_q, nq, n_d, dense)


@pytest.mark.parametrize("dim", [192, 256, 300, 320])
def test_wide_embeddings_k_pipelined_kernel(dim):
    """embedding dims above 128 (ColQwen3: 320) through the K-pipelined kernel."""
    dev = torch.device(DEV)
    g = torch.Generator().manual_seed(dim)
    lens = [int(x) for x in torch.randint(1, 700, (60,), generator=g)]
    q = torch.nn.functional.normalize(torch.randn(9, 32, dim, generator=g), dim=-1).bfloat16()
    bank = torch.nn.functional.normalize(torch.randn(sum(lens), dim, generator=g), dim=-1).bfloat16()
    ps = list(torch.split(bank, lens))
    got = cb.score_multi_vector(list(q), ps, device=dev)
    floors = O.reference_floors(lens, 128)
    lens_t = torch.tensor(lens, device=dev)
    pad = torch.nn.utils.rnn.pad_sequence([p.to(dev).float() for p in ps], batch_first=True)
    sim = torch.einsum("bnd,csd->bcns", q.to(dev).float(), pad)
    valid = torch.arange(pad.shape[1], device=dev)[None, :] < lens_t[:, None]
    mx = sim.masked_fill(~valid[None, :, None, :], float("-inf")).amax(dim=3)
    mx = torch.maximum(mx, torch.tensor(floors, device=dev)[None, :, None])
    want = mx.sum(dim=2).cpu()
    assert torch.allclose(got, want, rtol=1e-5, atol=2e-4)
    dense = torch.nn.functional.normalize(torch.randn(40, 515, dim, generator=g), dim=-1).bfloat16()
    got_d = cb.score_multi_vector(q, dense, device=dev)
    want_d = torch.einsum("bnd,csd->bcns", q.to(dev).float(), dense.to(dev).float()).amax(3).sum(2).cpu()
    assert torch.allclose(got_d, want_d, rtol=1e-5, atol=2e-4)


def test_wide_dim320_against_reference_golden():
    """K-pipelined scorer at ColQwen3's dim against the reference's own outputs (ragged, N_q up to 32)."""
    g = load_golden("wide_dim320.npz")
    qs = split_rows(from_bits(g["s_q"], (-1, 320)), g["s_qlen"])
    ps = split_rows(from_bits(g["s_p"], (-1, 320)), g["s_plen"])
    got = cb.score_multi_vector(qs, ps, device=DEV)
    want = torch.from_numpy(g["s_fp32"])
    assert rel_err(got, want) < 1e-4
    assert torch.equal(got.argmax(1), want.argmax(1))
    got_bf = cb.score_multi_vector(qs, ps, device=DEV, round_bf16=True)
    want_bf = torch.from_numpy(g["s_bf16"])
    ulp = want_bf.abs().clamp_min(1e-3) * 2.0 ** -7
    assert ((got_bf - want_bf).abs() <= ulp).all()


def _fp32_scores_chunked(q, docs, chunk=50):
    """Plain PyTorch fp32 MaxSim of a dense [n, L, D] bank on the GPU, chunked over documents (no TF32)."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        nq, nt, dim = q.shape
        q2 = q.float().reshape(nq * nt, dim)
        out = torch.empty(nq, docs.shape[0], dtype=torch.float32, device=q.device)
        for lo in range(0, docs.shape[0], chunk):
            d = docs[lo:lo + chunk].float()
            s = q2 @ d.reshape(-1, dim).t()
            out[:, lo:lo + chunk] = s.view(nq, nt, d.shape[0], d.shape[1]).amax(3).sum(1)
        return out
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


@pytest.mark.parametrize("n_docs", [3000])
def test_cfg4_geometry_128_queries_planted_positives(n_docs):
    """BASELINE configs[3] geometry on one GPU: 128 queries (32 query tiles -> 16 groups, group_sets = 8, 9 balanced
    partitions) x a dense 1030-token bank with one planted positive per query, against fp32 torch on the WHOLE matrix
    (processing_utils.py:179 in fp32): scores, argmax(1) and the planted top-1.  Round 1's cfg4 script reported
    planted_top1 = 1/128 at this geometry; no parity test reached it."""
    dev = torch.device(DEV)
    g = torch.Generator(device=dev).manual_seed(4)
    q = torch.nn.functional.normalize(torch.randn(128, 32, 128, device=dev, generator=g), dim=-1).bfloat16()
    docs = torch.empty(n_docs, 1030, 128, dtype=torch.bfloat16, device=dev)
    for lo in range(0, n_docs, 500):
        hi = min(lo + 500, n_docs)
        docs[lo:hi] = torch.nn.functional.normalize(torch.randn(hi - lo, 1030, 128, device=dev, generator=g), dim=-1).bfloat16()
    noise = torch.randn(128, 32, 128, device=dev, generator=g)
    planted = torch.tensor([(i * 97 + 13) % n_docs for i in range(128)], device=dev)
    assert planted.unique().numel() == 128
    for i in range(128):
        docs[planted[i], -32:] = torch.nn.functional.normalize(q[i].float() + 0.08 * noise[i], dim=-1).bfloat16()
    want = _fp32_scores_chunked(q, docs)
    assert torch.equal(want.argmax(1), planted)  # the recipe itself: the positive wins under the fp32 scorer
    bank = cb.DocBank.from_passages(docs, dev)
    assert bank.uniform_len == 1030 and bank.contiguous
    got = cb.maxsim(cb.QueryBlock(q, dev), bank)
    assert rel_err(got, want) < 1e-5
    assert torch.equal(got.argmax(1), planted)
    assert torch.equal(got.topk(10, dim=1).indices, want.topk(10, dim=1).indices)
    # the API path (host scores) and the other launch geometries give the same bits
    assert torch.equal(cb.score_multi_vector(q, bank), got.cpu())
    from colpali_b200 import _lib
    for opt, val in (("balanced", 0), ("cluster", 1), ("qtiles_per_cta", 1)):
        _lib.set_option(opt, val)
        try:
            assert torch.equal(cb.maxsim(cb.QueryBlock(q, dev), bank), got), opt
        finally:
            _lib.set_option(opt, 1 if opt == "balanced" else 0)


def test_cuda_graph_replays_with_new_queries():
    """A captured launch repeats with the same per-launch epoch; the partition-boundary exchange (balanced mode) must
    still hand over THIS replay's partials: replays with new queries in the static buffer equal eager launches."""
    dev = torch.device(DEV)
    g = torch.Generator().manual_seed(77)
    docs = torch.randn(60, 1030, 128, generator=g).bfloat16().to(dev)
    bank = cb.DocBank.from_passages(docs, dev)
    q_static = torch.randn(16, 32, 128, generator=g).bfloat16().to(dev)
    qb = cb.QueryBlock(q_static, dev)
    assert qb.flat.data_ptr() == q_static.data_ptr()  # zero-copy: replays see what is written into q_static
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        cb.maxsim(qb, bank)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out_a = cb.maxsim(qb, bank)
        out_b = cb.maxsim(qb, bank, want_argmax=True)[0]
    for trial in range(4):
        q_new = torch.randn(16, 32, 128, generator=g).bfloat16().to(dev)
        q_static.copy_(q_new)
        graph.replay()
        torch.cuda.synchronize()
        got_a, got_b = out_a.clone(), out_b.clone()
        want = cb.maxsim(cb.QueryBlock(q_new, dev), bank)
        torch.cuda.synchronize()
        assert torch.equal(got_a, want) and torch.equal(got_b, want), trial
