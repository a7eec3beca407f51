"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

A plain-PyTorch (CPU, fp32 or fp64) restatement of the ModalTune-GigaPath fine-tuning hot path of
martellab-sri/ModalTune, written from the behaviour of the reference, each function citing the reference file:line
it follows (paths relative to the reference root; ``TS/`` = ``models/prov_gigapath/gigapath/torchscale/``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module, and only as the checker / the timed CPU baseline.  ``modaltune_b200`` never imports it.

Parity pinning: the reference ships no tests or golden vectors, so the pin is the reference itself,
run in the build container: ``tests/golden/validate_oracle.py`` executes the unmodified reference modules (with the
import shims of ``tests/golden/ref_shims.py``) next to this file on identical weights/inputs and records the
agreement in ``tests/golden/ORACLE_VALIDATION.json``; ``tests/golden/make_golden.py`` stores reference outputs as
fixtures under ``tests/golden/*.pt`` which ``tests/test_oracle_golden.py`` replays against this file on every run.

Everything is functional over a ``state_dict``-style mapping ``sd`` that uses the REFERENCE's parameter names
(SURVEY.md §8b), so the same weights drive the reference, this oracle and the CUDA path.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# ---------------------------------------------------------------------------------------------------------------------
# Configuration of the path (model_configs/modaltune_gigapath_config.json, TS/model/LongNetConfig.py:166-179)
# ---------------------------------------------------------------------------------------------------------------------
EMBED = 768
HEADS = 16
HEAD_DIM = 48
FFN = 3072
DEPTH = 12
LN_EPS = 1e-5           # TS/architecture/config.py:40 (layernorm_eps)
ADAPTER_HEADS = 12
ADAPTER_DIM = 192       # int(768 * cffn_ratio 0.25), models/vitadapter/adapter_modules.py:153-155
DILATED_RATIO = (1, 2, 4, 8, 16)
INTERACTION_INDEXES = ((0, 3), (4, 7), (8, 11))


def optimal_segment_lengths(max_wsi_size: int = 262144, tile_size: int = 256) -> List[int]:
    """models/prov_gigapath/gigapath/slide_encoder.py:163-182 -> [1024, 5792, 32768, 185363, 1048576]."""
    max_seq_len = (max_wsi_size // tile_size) ** 2
    seg = np.linspace(np.log2(1024), int(np.log2(max_seq_len)), 5)
    return [int(v) for v in np.power(2, seg).astype(int)]


# ---------------------------------------------------------------------------------------------------------------------
# A0: positional table, patch embedding, cls assembly
# ---------------------------------------------------------------------------------------------------------------------
def sincos_table(ngrids: int = 1000, dim: int = EMBED) -> Tensor:
    """1-D factor of the 2-D sincos table (models/prov_gigapath/gigapath/pos_embed.py:34-81).

    ``pos_embed[1 + i*G + j] = [T[j] | T[i]]`` with ``T[p] = [sin(p*w) | cos(p*w)]``, ``w_k = 10000**(-k/(dim/4))``;
    float64 math, cast to fp32 like ``slide_encoder.py:149`` does.  Row 0 (cls) of ``pos_embed`` is all zeros.
    """
    quarter = dim // 4
    omega = 1.0 / 10000 ** (np.arange(quarter, dtype=np.float64) / quarter)
    out = np.einsum("m,d->md", np.arange(ngrids, dtype=np.float64), omega)
    return torch.from_numpy(np.concatenate([np.sin(out), np.cos(out)], axis=1)).float()  # [G, dim/2]


def coords_to_grid(coords: Tensor, ngrids: int = 1000) -> Tuple[Tensor, Tensor]:
    """slide_encoder.py:198-211: ``pos = floor(c0/256)*G + floor(c1/256) + 1`` -> (i, j) grid indices."""
    c = torch.floor(coords / 256.0)
    return c[..., 0].long(), c[..., 1].long()


def embed_tokens(sd: Dict[str, Tensor], feats: Tensor, coords: Tensor, table: Optional[Tensor] = None) -> Tensor:
    """PatchEmbed + pos + cls (slide_encoder.py:52-56, longvit_adapter.py:232-246, TS/architecture/encoder.py:342-385).

    feats [L, C], coords [L, 2] -> x [L+1, 768].  ``prepare_forward`` multiplies by embed_scale (=1, no_scale_embedding)
    and by (1 - padding_mask) (=1) and applies dropout (eval: identity).
    """
    dt = feats.dtype
    if table is None:
        table = sincos_table()
    x = F.linear(feats, sd["patch_embed.proj.weight"].to(dt), sd["patch_embed.proj.bias"].to(dt))
    i, j = coords_to_grid(coords)
    pos = torch.cat([table[j], table[i]], dim=-1).to(dt)
    x = x + pos
    cls = sd["cls_token"].reshape(1, -1).to(dt)  # + pos_embed[0] == 0
    return torch.cat([cls, x], dim=0)


# ---------------------------------------------------------------------------------------------------------------------
# A3-A5: LongNet dilated attention (TS/component/dilated_attention.py:22-144,146-262)
# ---------------------------------------------------------------------------------------------------------------------
def branch_geometry(n_tokens: int, segment_length: int, ratio: int, heads: int = HEADS):
    """Geometry of one (segment, dilation) branch.

    gathering() (:82-111): ``g = min(sl, N)``; N is zero padded to a multiple of g and cut into ``n_seg`` segments;
    dense_to_sparse() (:22-37): g is zero padded to a multiple of r and head h keeps local offsets ``o_h + j*r`` with
    ``o_h = floor(h*r/heads)``, ``j < m = ceil(g/r)``.
    """
    g = min(segment_length, n_tokens)
    n_seg = -(-n_tokens // g)
    m = -(-g // ratio)
    if n_seg > 1 and g % ratio != 0:
        # the reference itself mis-aligns positions on scatter here (:124) -- only reachable for N > 185363
        raise NotImplementedError("segment length not divisible by dilation with several segments")
    offsets = [(h * ratio) // heads for h in range(heads)]
    return g, n_seg, m, offsets


class _ScatterRows(torch.autograd.Function):
    """Write per-(segment, offset) blocks [hpb, c, d] into the strided rows of a dense [N, H, d] tensor (the reference's
    sparse_to_dense, dilated_attention.py:39-59) with a cheap strided-slice backward."""

    @staticmethod
    def forward(ctx, base, r, hpb, where, *blocks):
        out = base.clone()
        for (lo, hi, o), blk in zip(where, blocks):
            out[lo:hi:r, o * hpb:(o + 1) * hpb] = blk.transpose(0, 1)
        ctx.meta = (r, hpb, where)
        return out

    @staticmethod
    def backward(ctx, grad):
        r, hpb, where = ctx.meta
        return (None, None, None, None) + tuple(grad[lo:hi:r, o * hpb:(o + 1) * hpb].transpose(0, 1)
                                                for lo, hi, o in where)


def dilated_attention_core(q: Tensor, k: Tensor, v: Tensor, segment_lengths: Sequence[int],
                           ratios: Sequence[int], return_branches: bool = False):
    """q, k, v [N, H, d] -> merged attention [N, H*d].

    Per head h and branch b=(sl, r): slot j of segment s is position ``s*g + o_h + j*r``; a slot whose position is
    >= min(N, (s+1)*g) is a ZERO query/key (k = v = 0 -> score 0, value 0, counted in the softmax denominator, because
    the reference zero-pads after projection and flash-attn gets no mask; multihead_attention.py:109-119).  Branch
    outputs are mixed with ``softmax_b(lse_b)`` computed under no_grad (scattering(), :132-137); (head, position) pairs
    a branch does not own get weight 0 (lse = -1e8, :52).
    """
    N, H, d = q.shape
    scale = 1.0 / math.sqrt(d)
    dt = q.dtype
    outs, lses = [], []
    for sl, r in zip(segment_lengths, ratios):
        g, n_seg, m, offsets = branch_geometry(N, int(sl), int(r), H)
        hpb = H // r  # heads h with floor(h*r/H) == o are the contiguous block [o*hpb, (o+1)*hpb)
        o_parts = [[None] * r for _ in range(n_seg)]
        l_parts = [[None] * r for _ in range(n_seg)]
        for s in range(n_seg):
            seg_end = min(N, (s + 1) * g)
            for o in range(r):
                lo = s * g + o
                # slots j = 0..m-1 sit at positions lo + j*r; the real ones are the strided slice below
                qs = q[lo:seg_end:r, o * hpb:(o + 1) * hpb].transpose(0, 1)   # [hpb, c, d]
                c = qs.shape[1]
                if c == 0:
                    continue
                ks = k[lo:seg_end:r, o * hpb:(o + 1) * hpb].transpose(0, 1)
                vs = v[lo:seg_end:r, o * hpb:(o + 1) * hpb].transpose(0, 1)
                sc = torch.matmul(qs, ks.transpose(1, 2)) * scale             # [hpb, c, c]
                n_zero = m - c
                mx = sc.max(dim=-1, keepdim=True).values.detach()
                if n_zero > 0:
                    mx = mx.clamp(min=0.0)
                e = torch.exp(sc - mx)
                den = e.sum(-1, keepdim=True)
                if n_zero > 0:
                    den = den + n_zero * torch.exp(-mx)                        # zero keys: score 0, value 0
                o_parts[s][o] = (torch.matmul(e, vs) / den, lo, seg_end)
                l_parts[s][o] = (mx + torch.log(den)).squeeze(-1).detach()
        # assemble the dense per-branch tensors (zeros / -1e8 where a head does not own a position, :39-59)
        o_b = torch.zeros(N, H, d, dtype=dt)
        lse_b = torch.full((N, H), -1e8, dtype=dt)
        pieces = []
        for s in range(n_seg):
            for o in range(r):
                if o_parts[s][o] is None:
                    continue
                ob, lo, hi = o_parts[s][o]
                pieces.append((ob, l_parts[s][o], lo, hi, o))
        if q.requires_grad or k.requires_grad or v.requires_grad:
            o_b = _ScatterRows.apply(o_b, r, hpb, [(lo, hi, o) for _, _, lo, hi, o in pieces], *[p[0] for p in pieces])
        else:
            for ob, _, lo, hi, o in pieces:
                o_b[lo:hi:r, o * hpb:(o + 1) * hpb] = ob.transpose(0, 1)
        for _, lb, lo, hi, o in pieces:
            lse_b[lo:hi:r, o * hpb:(o + 1) * hpb] = lb.transpose(0, 1)
        outs.append(o_b)
        lses.append(lse_b)
    with torch.no_grad():
        L = torch.stack(lses, 0)
        w = torch.softmax(L, dim=0)
        w = torch.where(L <= -1e7, torch.zeros_like(w), w)
    out = 0
    for wb, ob in zip(w, outs):
        out = out + ob * wb[..., None].to(dt)
    out = out.reshape(N, H * d)
    if return_branches:
        return out, outs, lses
    return out


def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = LN_EPS) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), w.to(x.dtype), b.to(x.dtype), eps)


def dilated_self_attention(sd: Dict[str, Tensor], prefix: str, x: Tensor, segment_lengths, ratios) -> Tensor:
    """DilatedAttention.forward (dilated_attention.py:146-262): q/k/v proj, dilated core, inner_attn_ln, out_proj."""
    dt = x.dtype
    N = x.shape[0]
    lin = lambda name, t: F.linear(t, sd[f"{prefix}.{name}.weight"].to(dt), sd[f"{prefix}.{name}.bias"].to(dt))
    q = lin("q_proj", x).reshape(N, HEADS, HEAD_DIM)
    k = lin("k_proj", x).reshape(N, HEADS, HEAD_DIM)
    v = lin("v_proj", x).reshape(N, HEADS, HEAD_DIM)
    a = dilated_attention_core(q, k, v, segment_lengths, ratios)
    a = layer_norm(a, sd[f"{prefix}.inner_attn_ln.weight"], sd[f"{prefix}.inner_attn_ln.bias"])
    return lin("out_proj", a)


# ---------------------------------------------------------------------------------------------------------------------
# A1, A6: encoder layer (TS/architecture/encoder.py:121-175, TS/component/feedforward_network.py:132-143), eval mode
# ---------------------------------------------------------------------------------------------------------------------
def feed_forward(sd: Dict[str, Tensor], prefix: str, x: Tensor) -> Tensor:
    dt = x.dtype
    h = F.linear(x, sd[f"{prefix}.fc1.weight"].to(dt), sd[f"{prefix}.fc1.bias"].to(dt))
    h = F.gelu(h.float()).type_as(h)  # GELU is evaluated in fp32 whatever the input dtype (:136)
    h = layer_norm(h, sd[f"{prefix}.ffn_layernorm.weight"], sd[f"{prefix}.ffn_layernorm.bias"])
    return F.linear(h, sd[f"{prefix}.fc2.weight"].to(dt), sd[f"{prefix}.fc2.bias"].to(dt))


def encoder_layer(sd: Dict[str, Tensor], l: int, x: Tensor, segment_lengths, ratios, masks=None) -> Tensor:
    """Pre-LN block with sub-LN, alpha = 1 (encoder.py:137-175).  Dropout / DropPath are identity in eval; in train mode
    the reference applies Dropout(p) then DropPath to each residual branch before the add (encoder.py:149-155, 169-172;
    the FFN's own output dropout, feedforward_network.py:142, multiplies the same tensor): ``masks`` = the two
    multiplicative factors (keep / (1 - p) * path_scale, [N, 768]) of the attention and FFN branches."""
    p = f"encoder.layers.{l}"
    h = layer_norm(x, sd[f"{p}.self_attn_layer_norm.weight"], sd[f"{p}.self_attn_layer_norm.bias"])
    a = dilated_self_attention(sd, f"{p}.self_attn", h, segment_lengths, ratios)
    x = x + (a if masks is None else a * masks[0].to(a.dtype))
    h = layer_norm(x, sd[f"{p}.final_layer_norm.weight"], sd[f"{p}.final_layer_norm.bias"])
    f = feed_forward(sd, f"{p}.ffn", h)
    return x + (f if masks is None else f * masks[1].to(f.dtype))


# ---------------------------------------------------------------------------------------------------------------------
# A7-A9: Modal Adapter blocks (models/vitadapter/adapter_modules.py)
# ---------------------------------------------------------------------------------------------------------------------
def _mha(sd: Dict[str, Tensor], prefix: str, query: Tensor, key: Tensor, value: Tensor, heads: int = ADAPTER_HEADS):
    """torch.nn.MultiheadAttention(192, 12, kdim=vdim=768, batch_first=True), separate q/k/v weights, no masks."""
    dt = query.dtype
    E = query.shape[-1]
    b = sd[f"{prefix}.in_proj_bias"].to(dt)
    q = F.linear(query, sd[f"{prefix}.q_proj_weight"].to(dt), b[:E])
    k = F.linear(key, sd[f"{prefix}.k_proj_weight"].to(dt), b[E:2 * E])
    v = F.linear(value, sd[f"{prefix}.v_proj_weight"].to(dt), b[2 * E:])
    Lq, Lk, hd = q.shape[0], k.shape[0], E // heads
    q = q.reshape(Lq, heads, hd).transpose(0, 1)
    k = k.reshape(Lk, heads, hd).transpose(0, 1)
    v = v.reshape(Lk, heads, hd).transpose(0, 1)
    p = torch.softmax((q @ k.transpose(1, 2)) / math.sqrt(hd), dim=-1)
    o = (p @ v).transpose(0, 1).reshape(Lq, E)
    return F.linear(o, sd[f"{prefix}.out_proj.weight"].to(dt), sd[f"{prefix}.out_proj.bias"].to(dt))


def cross_attention_pre(sd, prefix, tgt, memory, pos=None, query_pos=None):
    """CrossAttentionLayer.forward_pre (adapter_modules.py:210-234), with_cffn=True.  Returns tgt + attn."""
    dt = tgt.dtype
    t2 = layer_norm(tgt, sd[f"{prefix}.norm.weight"], sd[f"{prefix}.norm.bias"])
    mem = layer_norm(memory, sd[f"{prefix}.norm_kq.weight"], sd[f"{prefix}.norm_kq.bias"])
    qin = t2 if query_pos is None else t2 + query_pos.to(dt)
    query = F.linear(qin, sd[f"{prefix}.q_proj.weight"].to(dt), sd[f"{prefix}.q_proj.bias"].to(dt))
    kv = mem if pos is None else mem + pos.to(dt)  # pos goes into the VALUE input too (:225-229)
    a = _mha(sd, f"{prefix}.multihead_attn", query, kv, kv)
    a = F.linear(a, sd[f"{prefix}.output_proj.weight"].to(dt), sd[f"{prefix}.output_proj.bias"].to(dt))
    return tgt + a


def injector(sd, prefix, x, c, pos):
    """Injector.forward (:359-369): x <- x + gamma * (x + attn(x, c, pos=gene_pe))."""
    a = cross_attention_pre(sd, f"{prefix}.attn", x, c, pos=pos, query_pos=None)
    return x + sd[f"{prefix}.gamma"].to(x.dtype) * a


def extractor(sd, prefix, c, x, pos):
    """Extractor.forward (:321-335): c <- c + (c + attn(c, x, query_pos=gene_pe)); c <- c + FFN_pre(c) (eval)."""
    a = cross_attention_pre(sd, f"{prefix}.attn", c, x, pos=None, query_pos=pos)
    c = c + a
    dt = c.dtype
    h = layer_norm(c, sd[f"{prefix}.ffn.norm.weight"], sd[f"{prefix}.ffn.norm.bias"])
    h = F.relu(F.linear(h, sd[f"{prefix}.ffn.linear1.weight"].to(dt), sd[f"{prefix}.ffn.linear1.bias"].to(dt)))
    h = F.linear(h, sd[f"{prefix}.ffn.linear2.weight"].to(dt), sd[f"{prefix}.ffn.linear2.bias"].to(dt))
    return c + h


def prompt_self_attention(sd, prefix, tgt, query_pos):
    """SelfAttentionLayer.forward_pre (:81-94), with_cffn=True: value is the un-positioned LN output."""
    dt = tgt.dtype
    t2 = layer_norm(tgt, sd[f"{prefix}.norm.weight"], sd[f"{prefix}.norm.bias"])
    qk = t2 + query_pos.to(dt)
    query = F.linear(qk, sd[f"{prefix}.q_proj.weight"].to(dt), sd[f"{prefix}.q_proj.bias"].to(dt))
    a = _mha(sd, f"{prefix}.self_attn", query, qk, t2)
    return tgt + F.linear(a, sd[f"{prefix}.output_proj.weight"].to(dt), sd[f"{prefix}.output_proj.bias"].to(dt))


# ---------------------------------------------------------------------------------------------------------------------
# Gene encoder (models/genomic_utils/gene_encoder.py:98-223), eval mode.  Produces the modal tokens; not a kernel target.
# ---------------------------------------------------------------------------------------------------------------------
def gene_encoder(sd, prefix, genes: Sequence[Tensor], depth: int = 3) -> Tensor:
    dt = genes[0].dtype
    W = lambda n: sd[f"{prefix}.{n}"].to(dt)
    rows = []
    for i, gi in enumerate(genes):  # gene_encode (:194-203): two SNN blocks (Linear+ELU) per pathway
        h = F.elu(F.linear(gi, W(f"gene_networks.{i}.0.0.weight"), W(f"gene_networks.{i}.0.0.bias")))
        h = F.elu(F.linear(h, W(f"gene_networks.{i}.1.0.weight"), W(f"gene_networks.{i}.1.0.bias")))
        rows.append(h)
    x = torch.cat(rows, 0)  # [G, 256]
    for dpt in range(depth):  # MLP-Mixer (:127-147): token mixing (Conv1d k=1 over groups) then channel mixing
        p = f"mlp_mixer.{dpt}"
        h = layer_norm(x, W(f"{p}.0.norm.weight"), W(f"{p}.0.norm.bias"))
        h = F.linear(h.t(), W(f"{p}.0.fn.0.weight").squeeze(-1), W(f"{p}.0.fn.0.bias"))
        h = F.gelu(h)
        h = F.linear(h, W(f"{p}.0.fn.3.weight").squeeze(-1), W(f"{p}.0.fn.3.bias")).t()
        x = x + h
        h = layer_norm(x, W(f"{p}.1.norm.weight"), W(f"{p}.1.norm.bias"))
        h = F.gelu(F.linear(h, W(f"{p}.1.fn.0.weight"), W(f"{p}.1.fn.0.bias")))
        h = F.linear(h, W(f"{p}.1.fn.3.weight"), W(f"{p}.1.fn.3.bias"))
        x = x + h
    x = layer_norm(x, W(f"mlp_mixer.{depth}.weight"), W(f"mlp_mixer.{depth}.bias"))
    x = F.linear(x, W(f"mlp_mixer.{depth + 1}.weight"), W(f"mlp_mixer.{depth + 1}.bias"))  # [G, 768]
    x = F.linear(x.t(), W("pathway_compression.weight"), W("pathway_compression.bias")).t()  # [64, 768] (:211-212)
    return x


# ---------------------------------------------------------------------------------------------------------------------
# A10: the adapter forward (models/aggregators/longvit_adapter.py:514-672 clinical, :205-347 gene-only)
# ---------------------------------------------------------------------------------------------------------------------
def adapter_forward(sd: Dict[str, Tensor], feats: Tensor, coords: Tensor, genes: Sequence[Tensor],
                    clinical: Optional[Tensor], task_token: Tensor, segment_lengths=None,
                    ratios=DILATED_RATIO, table: Optional[Tensor] = None, checkpoint_layers: bool = False) -> Tensor:
    """feats [L,1536], coords [L,2], genes list of [1,n_i], clinical [1,5] or None, task_token [k] -> [1, 256].

    prompt_agg='avg', token_agg='sum', use_prompt_sa, use_extra_extractor, multi_task>1 (the shipped JSON config).
    Note the adapter never applies ``self.norm`` / ``encoder.layer_norm`` (longvit_adapter.py:309-312).

    ``checkpoint_layers``: recompute each encoder layer in the backward (torch.utils.checkpoint, the reference's own
    ``checkpoint_activations`` switch, TS/architecture/encoder.py:317-319) instead of keeping the dense attention
    probabilities of all 36 layer passes alive -- ~5 GB per layer at 10k tiles, which no host holds for a whole step.
    Same arithmetic, same results; used by bench.py's CPU arm at the bench size.
    """
    if segment_lengths is None:
        segment_lengths = optimal_segment_lengths()
    dt = feats.dtype
    x = embed_tokens(sd, feats, coords, table)
    c = gene_encoder(sd, "gene_encoder", genes)                                           # [64, 768]
    t = F.linear(task_token.to(dt)[None], sd["task_weight.0.weight"].to(dt), sd["task_weight.0.bias"].to(dt))
    t = layer_norm(t, sd["task_weight.1.weight"], sd["task_weight.1.bias"])
    c = torch.cat([t, c], 0)                                                               # task token first (:576-580)
    if clinical is not None:
        h = F.relu(F.linear(clinical.to(dt), sd["clinical_mlp.0.weight"].to(dt), sd["clinical_mlp.0.bias"].to(dt)))
        h = F.linear(h, sd["clinical_mlp.2.weight"].to(dt), sd["clinical_mlp.2.bias"].to(dt))
        h = layer_norm(h, sd["clinical_mlp.3.weight"], sd["clinical_mlp.3.bias"])
        c = torch.cat([h, c], 0)                                                           # clinical token first (:583-584)
    pe = sd["gene_pe"].to(dt)
    cls, x = x[:1], x[1:]
    for i, (lo, hi) in enumerate(INTERACTION_INDEXES):
        if i > 0:
            c = prompt_self_attention(sd, f"prompt_selfattention.{i}", c, pe)
        x = injector(sd, f"interactions.{i}.injector", x, c, pe)
        x = torch.cat([cls, x], 0)
        for l in range(lo, hi + 1):
            if checkpoint_layers and torch.is_grad_enabled():
                from torch.utils.checkpoint import checkpoint
                x = checkpoint(encoder_layer, sd, l, x, segment_lengths, ratios, use_reentrant=False)
            else:
                x = encoder_layer(sd, l, x, segment_lengths, ratios)
        cls, x = x[:1], x[1:]
        c = extractor(sd, f"interactions.{i}.extractor", c, x, pe)
        if i == len(INTERACTION_INDEXES) - 1:
            for e in range(2):
                c = extractor(sd, f"interactions.{i}.extra_extractors.{e}", c, x, pe)
    n_lead = 2 if clinical is not None else 1
    gene_out = c[n_lead:].mean(0, keepdim=True)
    out = cls + gene_out + c[n_lead - 1:n_lead]                                            # + task token
    if clinical is not None:
        out = out + c[0:1]
    out = layer_norm(out, sd["final_norm.weight"], sd["final_norm.bias"])
    return F.linear(out, sd["final_project.weight"].to(dt), sd["final_project.bias"].to(dt))


# ---------------------------------------------------------------------------------------------------------------------
# Loss of one training step (train_modaltune.py:44-59, 156-179, 211-233)
# ---------------------------------------------------------------------------------------------------------------------
def text_targets(proj_sd: Dict[str, Tensor], text: Tensor) -> Tensor:
    """Frozen random Projection_layer: Conv1x1(512->256) -> LN([256,1,1]) -> ReLU -> Conv1x1(256->256); L2 norm."""
    dt = text.dtype
    h = F.linear(text, proj_sd["conv1.0.weight"].to(dt).flatten(1), proj_sd["conv1.0.bias"].to(dt))
    h = F.layer_norm(h, (h.shape[-1],), proj_sd["conv1.1.weight"].to(dt).flatten(), proj_sd["conv1.1.bias"].to(dt).flatten(), 1e-5)
    h = F.relu(h)
    h = F.linear(h, proj_sd["conv1.3.weight"].to(dt).flatten(1), proj_sd["conv1.3.bias"].to(dt))
    return h / h.norm(dim=-1, keepdim=True)


def distill_loss(logits: Tensor, text_proj: Tensor) -> Tensor:
    """10 * KLDiv_sum(log_softmax(z/|z|), softmax(t[[0,1,3]])), temperature 1 (train_modaltune.py:225-233)."""
    z = logits / logits.norm(dim=-1, keepdim=True)
    return F.kl_div(F.log_softmax(z, dim=1), F.softmax(text_proj[[0, 1, 3]], dim=1), reduction="sum") * 10.0


def training_step(sd, proj_sd, feats, coords, genes, clinical, text, num_tasks: int = 3, segment_lengths=None,
                  table=None, checkpoint_layers: bool = False):
    """Three task passes + loss (multitask_forward, train_modaltune.py:156-179).  Returns (loss, logits [3,256])."""
    eye = torch.eye(num_tasks, dtype=feats.dtype)
    logits = torch.cat([adapter_forward(sd, feats, coords, genes, clinical, eye[t], segment_lengths, table=table,
                                        checkpoint_layers=checkpoint_layers) for t in range(3)], 0)
    return distill_loss(logits, text_targets(proj_sd, text)), logits
