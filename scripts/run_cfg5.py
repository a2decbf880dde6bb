"""BASELINE configs[4]: end-to-end pages/sec -- random-init ColQwen2-2B (the reference's model class over the HF Qwen2-VL
backbone) -> custom_text_proj head kernel -> fused MaxSim, synthetic 448 x 448 pages, data-parallel replicas.

    python scripts/run_cfg5.py [--pages 1000] [--batch 32]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/run_cfg5.py   (N replicas)

The model is ``colpali_engine.models.ColQwen2`` from baseline/_ref (the unmodified reference; no weights are downloaded:
``Qwen2VLConfig`` of Qwen2-VL-2B, random init, bf16).  Its ``forward`` tail (modeling_colqwen2.py:65-74) is replaced by
``colpali_b200.fused_head`` exactly as INTEGRATION.md shows; the backbone is stock HF code (library, not this repo's
product).  Inputs are synthesised directly in the processor's output format (no tokenizer files offline): one 448 x 448
page = a 32 x 32 patch grid -> 1024 patch rows of 3*2*14*14 values -> 256 merged visual tokens, wrapped in the 17-token
visual prompt of ColQwen2Processor (processing_colqwen2.py:22-24) -> 273 tokens per page.
Every rank embeds pages/world pages, keeps them in a device-resident DocBank and scores 32 queries against them
(pure replicas: no collective on the data path, SURVEY.md section 8e "head / backbone: replicas only").
Prints one JSON line on rank 0: pages/s (whole job, max over ranks) and the backbone / head / scorer split.
"""
import argparse, json, os, sys, time
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import colpali_b200 as cb

GRID, N_VIS, N_PROMPT = 32, 256, 17


def build_model(dev):
    from colpali_engine.models import ColQwen2
    from transformers.models.qwen2_vl import Qwen2VLConfig

    cfg = Qwen2VLConfig(  # Qwen2-VL-2B-Instruct
        text_config=dict(hidden_size=1536, intermediate_size=8960, num_hidden_layers=28, num_attention_heads=12,
                         num_key_value_heads=2, vocab_size=151936, max_position_embeddings=32768, rms_norm_eps=1e-6,
                         rope_parameters={"rope_type": "default", "rope_theta": 1000000.0, "mrope_section": [16, 24, 24]},
                         tie_word_embeddings=True),
        vision_config=dict(depth=32, embed_dim=1280, hidden_size=1536, num_heads=16, mlp_ratio=4, patch_size=14,
                           spatial_merge_size=2, temporal_patch_size=2, in_channels=3),
    )
    torch.manual_seed(0)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(dev):
            model = ColQwen2(cfg)
    finally:
        torch.set_default_dtype(prev)
    return model.eval(), cfg


def make_batch(n, cfg, dev, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    ids = torch.randint(1000, 50000, (n, N_PROMPT + N_VIS), device=dev, generator=g)
    ids[:, 3] = cfg.vision_start_token_id
    ids[:, 4:4 + N_VIS] = cfg.image_token_id
    ids[:, 4 + N_VIS] = cfg.vision_end_token_id
    pix = torch.randn(n, GRID * GRID, 3 * 2 * 14 * 14, device=dev, generator=g).bfloat16()  # seeded-noise "pages"
    thw = torch.tensor([[1, GRID, GRID]] * n, device=dev)
    # mm_token_type_ids: what the HF processor returns beside input_ids (1 = image token), needed for M-RoPE positions
    return dict(input_ids=ids, attention_mask=torch.ones_like(ids), pixel_values=pix, image_grid_thw=thw,
                mm_token_type_ids=(ids == cfg.image_token_id).long())


def backbone(model, batch):
    """ColQwen2.forward up to last_hidden_state (modeling_colqwen2.py:48-63), stock HF Qwen2VLModel."""
    from transformers.models.qwen2_vl import Qwen2VLModel

    kw = dict(batch)
    offsets = kw["image_grid_thw"][:, 1] * kw["image_grid_thw"][:, 2]
    kw["pixel_values"] = torch.cat([p[:o] for p, o in zip(kw["pixel_values"], offsets)], dim=0)
    return Qwen2VLModel.forward(model, **kw, use_cache=False, output_hidden_states=True, return_dict=True).last_hidden_state


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pages", type=int, default=1000)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--queries", type=int, default=32)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model, cfg = build_model(dev)
    n_params = sum(p.numel() for p in model.parameters())
    w, b = model.custom_text_proj.weight, model.custom_text_proj.bias
    my_pages = a.pages // world + (1 if rank < a.pages % world else 0)
    qs = torch.nn.functional.normalize(torch.randn(a.queries, 32, 128, device=dev), dim=-1).bfloat16()
    qb = cb.QueryBlock(qs, dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    with torch.no_grad():
        # ---- parity of the swapped tail on one batch: reference forward vs backbone + fused head ----------------------
        batch = make_batch(min(a.batch, 8), cfg, dev, 99)
        ref = model(**batch)
        h = backbone(model, batch)
        fused = cb.fused_head(h, w, b, batch["attention_mask"])
        same = float((fused == ref).float().mean())
        max_diff = float((fused.float() - ref.float()).abs().max())
        # warm-up
        for _ in range(2):
            cb.fused_head(backbone(model, make_batch(a.batch, cfg, dev, 1)), w, b, None)
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        t_back = t_head = 0.0
        embs = []
        e_all0, e_all1 = ev(), ev()
        e_all0.record()
        done = 0
        while done < my_pages:
            n = min(a.batch, my_pages - done)
            batch = make_batch(n, cfg, dev, 1000 * rank + done)
            e0, e1, e2 = ev(), ev(), ev()
            e0.record()
            h = backbone(model, batch)
            e1.record()
            emb = cb.fused_head(h, w, b, batch["attention_mask"])  # [n, 273, 128] bf16, unit rows
            e2.record()
            embs.append((emb, e0, e1, e2))
            done += n
        bank = cb.DocBank.from_passages(torch.cat([x[0] for x in embs], 0), dev)
        es0, es1 = ev(), ev()
        es0.record()
        scores = cb.maxsim(qb, bank)
        es1.record()
        top = scores.argmax(1).cpu()  # the device -> host read of the step's result
        e_all1.record()
        torch.cuda.synchronize()
        for _, e0, e1, e2 in embs:
            t_back += e0.elapsed_time(e1); t_head += e1.elapsed_time(e2)
        total_ms = torch.tensor([e_all0.elapsed_time(e_all1)], device=dev)
        if world > 1: dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        # scorer parity on the produced embeddings: fp32 torch on a document subset
        sub = torch.cat([x[0] for x in embs], 0)[:64].float()
        want = torch.einsum("bnd,csd->bcns", qs.float(), sub).amax(3).sum(2)
        rel = float(((scores[:, :64] - want).abs() / want.abs().clamp_min(1e-3)).max())
    if rank == 0:
        tokens = my_pages * (N_PROMPT + N_VIS)
        print(json.dumps({
            "config": "cfg5 end-to-end: random-init ColQwen2-2B (reference class, HF backbone) -> fused head -> fused MaxSim",
            "world": world, "pages": a.pages, "pages_per_rank": my_pages, "batch": a.batch, "tokens_per_page": N_PROMPT + N_VIS,
            "params": n_params, "pages_per_s": a.pages / float(total_ms) * 1e3, "total_ms": float(total_ms),
            "rank0_backbone_ms": t_back, "rank0_head_ms": t_head, "rank0_scorer_ms": es0.elapsed_time(es1),
            "head_tokens_per_s": tokens / t_head * 1e3, "head_share_of_step": t_head / float(total_ms),
            "head_bit_identical_to_reference_forward": same, "head_max_abs_diff": max_diff,
            "scorer_max_rel_err_vs_fp32_torch": rel, "queries": a.queries, "top1_ids_head": top[:4].tolist(),
        }), flush=True)
        assert same > 0.99 and max_diff < 0.01 and rel < 1e-4
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
