"""Fused multi-vector projection head (custom_text_proj -> L2 normalise -> masks) on B200.

Replaces the last lines of every ``Col*`` model ``forward`` in the reference, e.g.
``colpali_engine/models/qwen2/colqwen2/modeling_colqwen2.py:65-74``::

    proj = self.custom_text_proj(hidden_states)
    proj = proj / proj.norm(dim=-1, keepdim=True)
    proj = proj * kwargs["attention_mask"].unsqueeze(-1)
    [proj = proj * (input_ids == image_token_id).unsqueeze(-1)]

with one kernel (csrc/head_sm100.cu) that reads ``hidden_states`` once.  The module attribute stays an
``nn.Linear`` named ``custom_text_proj`` (state-dict keys and LoRA targeting are untouched); only the functional
tail is swapped -- see INTEGRATION.md for the three-line patch per model file.

Forward is the fused kernel.  Backward (training) recomputes the reference expression with library GEMMs under
autograd -- the head's backward is a plain GEMM pair and not on the serving hot path.
"""

from __future__ import annotations

from typing import Optional

import torch

from . import _lib

HEAD_DIM = 128      # the validated kernel (head_sm100.cu)
MAX_HEAD_DIM = 320  # multiples of 32 above 128 go through head_wide_sm100.cu (ColQwen3: 320)


def _supported_dim(dim: int) -> bool:
    return dim == HEAD_DIM or (HEAD_DIM < dim <= MAX_HEAD_DIM and dim % 32 == 0)


def _reference_tail(h, weight, bias, attention_mask, extra_mask, clamp_norm):
    proj = torch.nn.functional.linear(h, weight, bias)
    norm = proj.norm(dim=-1, keepdim=True)
    if clamp_norm:
        norm = norm.clamp_min(1e-12)
    proj = proj / norm
    if attention_mask is not None:
        proj = proj * attention_mask.unsqueeze(-1)
    if extra_mask is not None:
        proj = proj * extra_mask.unsqueeze(-1)
    return proj


class _FusedHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, weight, bias, attention_mask, extra_mask, clamp_norm, single_rounding):
        dev = h.device
        if dev.type != "cuda":
            raise _lib.ColpaliB200Error("fused_head needs CUDA tensors (sm_100a); there is no CPU path")
        dim = int(weight.shape[0])
        if not _supported_dim(dim):
            raise _lib.ColpaliB200Error(f"projection dim {dim} is not supported by this build "
                                        f"({HEAD_DIM}, or a multiple of 32 up to {MAX_HEAD_DIM})")
        lib = _lib.load()
        lead, hidden = h.shape[:-1], h.shape[-1]
        h2 = h.detach().reshape(-1, hidden)
        if h2.dtype != torch.bfloat16:
            h2 = h2.to(torch.bfloat16)
        h2 = h2.contiguous()
        w = weight.detach().to(torch.bfloat16).contiguous()
        b = bias.detach().to(torch.bfloat16).contiguous() if bias is not None else None
        n = h2.shape[0]
        am = attention_mask.reshape(-1).to(torch.int64).contiguous() if attention_mask is not None else None
        em = extra_mask.reshape(-1).to(torch.uint8).contiguous() if extra_mask is not None else None
        if am is not None and am.numel() != n:
            raise ValueError(f"attention_mask has {am.numel()} entries for {n} tokens")
        if em is not None and em.numel() != n:
            raise ValueError(f"extra mask has {em.numel()} entries for {n} tokens")
        out = torch.empty(n, dim, dtype=torch.bfloat16, device=dev)
        flags = (_lib.CPB_HEAD_CLAMP_NORM if clamp_norm else 0) | (_lib.CPB_HEAD_SINGLE_ROUNDING if single_rounding else 0)
        with torch.cuda.device(dev):
            rc = lib.cpb_head_fwd(
                h2.data_ptr(), n, hidden, w.data_ptr(), b.data_ptr() if b is not None else None, dim,
                am.data_ptr() if am is not None else None, em.data_ptr() if em is not None else None,
                out.data_ptr(), flags, torch.cuda.current_stream(dev).cuda_stream,
            )
        _lib.check(rc, "cpb_head_fwd")
        _lib.count_launches(1)
        ctx.save_for_backward(h, weight, bias if bias is not None else torch.empty(0, device=dev), attention_mask
                              if attention_mask is not None else torch.empty(0, device=dev),
                              extra_mask if extra_mask is not None else torch.empty(0, device=dev))
        ctx.flags = (bias is not None, attention_mask is not None, extra_mask is not None, clamp_norm)
        return out.view(*lead, dim).to(h.dtype) if h.dtype != torch.bfloat16 else out.view(*lead, dim)

    @staticmethod
    def backward(ctx, grad_out):
        h, weight, bias, am, em = ctx.saved_tensors
        has_b, has_am, has_em, clamp_norm = ctx.flags
        with torch.enable_grad():
            hh = h.detach().requires_grad_(ctx.needs_input_grad[0])
            ww = weight.detach().requires_grad_(ctx.needs_input_grad[1])
            bb = bias.detach().requires_grad_(ctx.needs_input_grad[2]) if has_b else None
            out = _reference_tail(hh, ww, bb, am if has_am else None, em if has_em else None, clamp_norm)
            wanted = [t for t in (hh, ww, bb) if t is not None and t.requires_grad]
            grads = torch.autograd.grad(out, wanted, grad_out.to(out.dtype), allow_unused=True) if wanted else ()
        it = iter(grads)
        gh = next(it) if hh.requires_grad else None
        gw = next(it) if ww.requires_grad else None
        gb = next(it) if (bb is not None and bb.requires_grad) else None
        return gh, gw, gb, None, None, None, None


def fused_head(hidden_states: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
               attention_mask: Optional[torch.Tensor] = None, image_mask: Optional[torch.Tensor] = None, *,
               clamp_norm: bool = False, single_rounding: bool = False) -> torch.Tensor:
    """``[..., hidden] -> [..., dim]`` unit-norm rows, zero rows where masked (dim = ``weight.shape[0]``).

    ``weight`` / ``bias`` are ``custom_text_proj``'s parameters; ``attention_mask`` and ``image_mask`` have the
    shape of ``hidden_states`` without the last dim.  By default the reference's three bf16 roundings are
    reproduced (bit-identical to the bf16 reference on > 99.9% of elements); ``single_rounding=True`` keeps fp32
    until the store (closer to the exact value, not to the reference).
    """
    return _FusedHeadFn.apply(hidden_states, weight, bias, attention_mask, image_mask, clamp_norm, single_rounding)
