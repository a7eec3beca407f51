"""Kernel-level parity on the B200: every C-ABI entry point against the oracle-based CPU statement of the same op
(``tests/cpu_kernels.py``) on identical seeded inputs.  fp32 kernels: tight tolerance; bf16 storage: bf16 rounding.
"""
import math

import pytest
import torch

from modaltune_b200 import _lib, ops
from modaltune_b200.slide_encoder import DILATED_RATIO, optimal_segment_lengths
from tests import cpu_kernels as C

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def tol(dtype):
    return 2e-5 if dtype == torch.float32 else 2e-2


def test_library_loads_on_sm100():
    lib = _lib.load()
    assert lib.mt_version() >= 100
    assert lib.mt_device_is_sm100() == 1


@pytest.mark.parametrize("cols", [768, 3072])
@pytest.mark.parametrize("xdt,ydt", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16),
                                     (torch.bfloat16, torch.bfloat16)])
@pytest.mark.parametrize("rows", [1, 66, 1037])
def test_layernorm_fwd_bwd(rows, cols, xdt, ydt):
    g = torch.Generator().manual_seed(rows * 7 + cols)
    x = (torch.randn(rows, cols, generator=g) * 1.7 + 0.3).to(xdt)
    gamma, beta = 1 + 0.1 * torch.randn(cols, generator=g), 0.1 * torch.randn(cols, generator=g)
    dy = torch.randn(rows, cols, generator=g).to(ydt)
    res = torch.randn(rows, cols, generator=g)
    y_c, m_c, r_c = C.layernorm_fwd(x, gamma, beta, ydt)
    y, m, r = ops.layernorm_fwd(x.to(DEV), gamma.to(DEV), beta.to(DEV), ydt)
    assert rel(y, y_c) < tol(ydt) and rel(m, m_c) < 1e-5 and rel(r, r_c) < 1e-4
    want_w = xdt == torch.float32 and cols == 768
    dx_c, dg_c, db_c = C.layernorm_bwd(dy, x, gamma, m_c, r_c, torch.float32, residual=res, want_wgrad=want_w)
    dx, dg, db = ops.layernorm_bwd(dy.to(DEV), x.to(DEV), gamma.to(DEV), m, r, torch.float32, residual=res.to(DEV),
                                   want_wgrad=want_w)
    assert rel(dx, dx_c) < 2e-5
    if want_w:
        assert rel(dg, dg_c) < 1e-4 and rel(db, db_c) < 1e-4
    # the bf16 twin written in the same pass is exactly the rounded fp32 result
    dx2, _, _ = ops.layernorm_bwd(dy.to(DEV), x.to(DEV), gamma.to(DEV), m, r, torch.float32, residual=res.to(DEV),
                                  bf16_twin=True)
    assert torch.equal(dx2, dx) and torch.equal(dx2._mt_bf16, dx.to(torch.bfloat16))


def test_layernorm_fused_add_and_add_layernorm():
    g = torch.Generator().manual_seed(5)
    x, a = torch.randn(66, 768, generator=g), torch.randn(66, 768, generator=g)
    gamma, beta = 1 + 0.1 * torch.randn(768, generator=g), 0.1 * torch.randn(768, generator=g)
    y_c, _, _ = C.layernorm_fwd(x, gamma, beta, torch.float32, add=a)
    y, _, _ = ops.layernorm_fwd(x.to(DEV), gamma.to(DEV), beta.to(DEV), torch.float32, add=a.to(DEV))
    assert rel(y, y_c) < 2e-5
    ab = torch.randn(768, generator=g)
    for adt in (torch.float32, torch.bfloat16):
        for bias in (None, ab):
            xo_c, y_c, m_c, r_c = C.add_layernorm_fwd(x, a.to(adt), gamma, beta, adt, abias=bias)
            xo, y, m, r = ops.add_layernorm_fwd(x.to(DEV), a.to(adt).to(DEV), gamma.to(DEV), beta.to(DEV), adt,
                                                abias=None if bias is None else bias.to(DEV))
            assert rel(xo, xo_c) < 1e-6 and rel(y, y_c) < tol(adt) and rel(m, m_c) < 1e-5 and rel(r, r_c) < 1e-4
        yb = ops.residual_bias_add(x.to(DEV), a.to(adt).to(DEV), ab.to(DEV))
        assert rel(yb, C.residual_bias_add(x, a.to(adt), ab)) < 1e-6


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_gelu_ln_fwd_bwd(dt):
    g = torch.Generator().manual_seed(11)
    h = (torch.randn(513, 3072, generator=g) * 1.5).to(dt)
    gamma, beta = 1 + 0.1 * torch.randn(3072, generator=g), 0.1 * torch.randn(3072, generator=g)
    dy = torch.randn(513, 3072, generator=g).to(dt)
    y_c, m_c, r_c = C.gelu_ln_fwd(h, gamma, beta, dt)
    y, m, r = ops.gelu_ln_fwd(h.to(DEV), gamma.to(DEV), beta.to(DEV), dt)
    assert rel(y, y_c) < tol(dt) and rel(m, m_c) < 1e-4 and rel(r, r_c) < 1e-4
    dh_c = C.gelu_ln_bwd(dy, h, gamma, m_c, r_c, dt)
    dh = ops.gelu_ln_bwd(dy.to(DEV), h.to(DEV), gamma.to(DEV), m, r, dt)
    assert rel(dh, dh_c) < tol(dt)
    hb = torch.randn(3072, generator=g)   # bias of the producing GEMM folded into the kernels
    y_c, m_c, r_c = C.gelu_ln_fwd(h, gamma, beta, dt, hbias=hb)
    y, m, r = ops.gelu_ln_fwd(h.to(DEV), gamma.to(DEV), beta.to(DEV), dt, hbias=hb.to(DEV))
    assert rel(y, y_c) < tol(dt) and rel(m, m_c) < 1e-4
    dh_c = C.gelu_ln_bwd(dy, h, gamma, m_c, r_c, dt, hbias=hb)
    dh = ops.gelu_ln_bwd(dy.to(DEV), h.to(DEV), gamma.to(DEV), m, r, dt, hbias=hb.to(DEV))
    assert rel(dh, dh_c) < tol(dt)


def test_embed_assemble_and_cast_and_gated_residual():
    g = torch.Generator().manual_seed(3)
    L = 777
    proj = torch.randn(L, 768, generator=g)
    bias, cls = torch.randn(768, generator=g), torch.randn(768, generator=g)
    from modaltune_b200.slide_encoder import sincos_factor
    table = sincos_factor(1000, 768)
    cells = torch.randint(0, 999, (L, 2), generator=g).float() * 256.0 + torch.rand(L, 2, generator=g) * 200
    x_c = C.embed_assemble(proj, bias, cells, table, cls)
    for dt in (torch.float32, torch.bfloat16):
        x = ops.embed_assemble(proj.to(dt).to(DEV), bias.to(DEV), cells.to(DEV), table.to(DEV), cls.to(DEV))
        assert rel(x, C.embed_assemble(proj.to(dt), bias, cells, table, cls)) < 1e-6
    assert rel(ops.embed_assemble(proj.to(DEV), bias.to(DEV), cells.to(DEV), table.to(DEV), cls.to(DEV)), x_c) < 1e-6
    v = torch.randn(100003, generator=g)
    assert torch.equal(ops.cast(v.to(DEV), torch.bfloat16).cpu(), v.to(torch.bfloat16))
    a, b, gate = torch.randn(301, 768, generator=g), torch.randn(301, 768, generator=g), torch.randn(768, generator=g)
    dy = torch.randn(301, 768, generator=g)
    for dt in (torch.float32, torch.bfloat16):
        y = ops.gated_residual_fwd(a.to(DEV), b.to(dt).to(DEV), gate.to(DEV))
        assert rel(y, C.gated_residual_fwd(a, b.to(dt), gate)) < 1e-6
        da, db, dg = ops.gated_residual_bwd(dy.to(DEV), a.to(DEV), b.to(dt).to(DEV), gate.to(DEV))
        da_c, db_c, dg_c = C.gated_residual_bwd(dy, a, b.to(dt), gate)
        assert rel(da, da_c) < 1e-6 and rel(db, db_c) < tol(dt) and rel(dg, dg_c) < 1e-4


def test_embed_assemble_follows_the_reference_flat_index():
    """pos = floor(c0/256)*G + floor(c1/256) + 1 into the flat [1 + G*G, E] table (slide_encoder.py:198-211): a column
    index >= G wraps into the next grid row exactly as the reference's gather does, and pos == 0 is the zero cls row."""
    from modaltune_b200.slide_encoder import sincos_factor
    G = 1000
    table = sincos_factor(G, 768)
    g = torch.Generator().manual_seed(9)
    cells = torch.tensor([[3, 1005], [0, 0], [998, 999], [7, 2500], [0, -1]], dtype=torch.float32)
    coords = cells * 256.0 + 17.0
    L = coords.shape[0]
    proj, bias, cls = torch.randn(L, 768, generator=g), torch.randn(768, generator=g), torch.randn(768, generator=g)
    pos = (torch.floor(coords[:, 0] / 256.0) * G + torch.floor(coords[:, 1] / 256.0)).long() + 1
    flat_rows = torch.zeros(L, 768)
    for n, p_ in enumerate(pos.tolist()):
        if p_ > 0:   # pos_embed[1 + i*G + j] = [T[j] | T[i]]
            flat_rows[n] = torch.cat([table[(p_ - 1) % G], table[(p_ - 1) // G]])
    want = torch.cat([cls[None], proj + bias + flat_rows], 0)
    got = ops.embed_assemble(proj.to(DEV), bias.to(DEV), coords.to(DEV), table.to(DEV), cls.to(DEV))
    assert rel(got, want) < 1e-6


@pytest.mark.parametrize("lq,lk", [(700, 66), (66, 700), (66, 10000), (5000, 13), (1, 1), (129, 65)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_cross_attention_core(lq, lk, dt):
    """impl 0: the exact fp32-math SIMT kernels (fp32 parity mode) against the oracle-built stand-in."""
    g = torch.Generator().manual_seed(lq * 31 + lk)
    q, k, v = (torch.randn(n, 192, generator=g).to(dt) for n in (lq, lk, lk))
    d_o = torch.randn(lq, 192, generator=g).to(dt)
    o_c, lse_c = C.cross_attn_fwd(q, k, v, 12)
    o, lse = ops.cross_attn_fwd(q.to(DEV), k.to(DEV), v.to(DEV), 12, impl=0)
    assert rel(o, o_c) < tol(dt) and rel(lse, lse_c) < 1e-5
    dq_c, dk_c, dv_c = C.cross_attn_bwd(q, k, v, o_c, d_o, lse_c, 12)
    dq, dk, dv = ops.cross_attn_bwd(q.to(DEV), k.to(DEV), v.to(DEV), o, d_o.to(DEV), lse, 12, impl=0)
    t = 5e-5 if dt == torch.float32 else 3e-2  # bf16: delta = dO.o uses the rounded o
    assert rel(dq, dq_c) < t and rel(dk, dk_c) < t and rel(dv, dv_c) < t


@pytest.mark.parametrize("lq,lk", [(700, 66), (66, 700), (66, 10000), (10000, 66), (5000, 13), (1, 1), (129, 65), (66, 66),
                                   (73, 145), (40000, 67)])
@pytest.mark.parametrize("packed", [False, True])
def test_cross_attention_tensor_core(lq, lk, packed):
    """impl 1 (the bf16-mode default): mma.sync TF32 operands / fp32 accumulation, against the fp64-accurate stand-in
    within TF32 rounding (2^-11 per operand), and the strided k | v layout of a fused projection buffer."""
    g = torch.Generator().manual_seed(lq * 31 + lk + 7)
    q, d_o = torch.randn(lq, 192, generator=g), torch.randn(lq, 192, generator=g)
    kv = torch.randn(lk, 384, generator=g)
    k, v = kv[:, :192].contiguous(), kv[:, 192:].contiguous()
    o_c, lse_c = C.cross_attn_fwd(q, k, v, 12)
    dq_c, dk_c, dv_c = C.cross_attn_bwd(q, k, v, o_c, d_o, lse_c, 12)
    if packed:
        kvd = kv.to(DEV)
        kd, vd = kvd[:, :192], kvd[:, 192:]
    else:
        kd, vd = k.to(DEV), v.to(DEV)
    qd = q.to(DEV)
    o, lse = ops.cross_attn_fwd(qd, kd, vd, 12, impl=1)
    assert rel(o, o_c) < 3e-3 and rel(lse, lse_c) < 1e-3
    dq, dk, dv = ops.cross_attn_bwd(qd, kd, vd, o, d_o.to(DEV), lse, 12, impl=1, packed_kv=packed)
    if packed:
        assert dk._base is dv._base and dk._base.shape == (lk, 384)
    # relative to the larger of the gradient and what feeds it: with a single key the exact dq / dk are zero (P = 1)
    # and TF32 leaves rounding noise of the size of d_o * 2^-11
    scale = lambda ref: max(float(ref.abs().max()), 0.1 * float(d_o.abs().max()))
    for got, ref in ((dq, dq_c), (dk, dk_c), (dv, dv_c)):
        assert float((got.cpu().double() - ref.double()).abs().max()) < 5e-3 * scale(ref)
    # and against the SIMT kernels on the same device inputs (tighter on the statistics)
    o0, lse0 = ops.cross_attn_fwd(qd, kd, vd, 12, impl=0)
    assert rel(lse, lse0) < 1e-3 and rel(o, o0) < 3e-3


@pytest.mark.parametrize("rows", [1, 37, 1200, 10001])
def test_layernorm_bwd_with_ffn_prep_of_the_layer_below(rows):
    """mt_layernorm_bwd_ffn_prep = mt_layernorm_bwd (+ residual, + bf16 twin) and mt_ffn_bwd_prep of the layer below in
    one kernel: against the two separate kernels and against the CPU stand-ins (oracle arithmetic)."""
    g = torch.Generator().manual_seed(rows)
    dy, x, res, x1b = (torch.randn(rows, 768, generator=g) for _ in range(4))
    gamma, c1, c2 = (torch.randn(768, generator=g) for _ in range(3))
    mean, mean_f = torch.randn(rows, generator=g) * 0.1, torch.randn(rows, generator=g)
    rstd, rstd_f = torch.rand(rows, generator=g) + 0.5, torch.rand(rows, generator=g) + 0.5
    d = lambda t: t.to(DEV)
    dx, twin, rowv = ops.layernorm_bwd_ffn_prep(d(dy), d(x), d(gamma), d(mean), d(rstd), d(res),
                                                (d(x1b), d(c1), d(c2), d(mean_f), d(rstd_f), 3072))
    dx_s, _, _ = ops.layernorm_bwd(d(dy), d(x), d(gamma), d(mean), d(rstd), torch.float32, residual=d(res), bf16_twin=True)
    rowv_s, twin_s = ops.ffn_bwd_prep(dx_s, d(x), d(x1b), d(c1), d(c2), d(mean_f), d(rstd_f), 3072)
    # the same arithmetic in two kernels: equal up to the order of the fused multiply-adds
    assert rel(dx, dx_s) < 1e-6 and rel(twin, twin_s) < 1e-2 and torch.equal(twin, dx.to(torch.bfloat16))
    assert rel(rowv[:, :2], rowv_s[:, :2]) < 1e-6 and rel(rowv[:, 2:], rowv_s[:, 2:]) < 5e-3
    dx_c, twin_c, rowv_c = C.layernorm_bwd_ffn_prep(dy, x, gamma, mean, rstd, res, (x1b, c1, c2, mean_f, rstd_f, 3072))
    assert rel(dx, dx_c) < 1e-5 and rel(rowv[:, :2], rowv_c[:, :2]) < 1e-6
    # m1, m2 are sums of 768 products of bf16-rounded gradients: one element rounding the other way moves them by 2^-9
    assert rel(rowv[:, 2:], rowv_c[:, 2:]) < 5e-3


@pytest.mark.parametrize("rows,cols", [(10000, 384), (10000, 192), (66, 768), (1, 4), (7, 3072), (40001, 768)])
def test_colsum(rows, cols):
    g = torch.Generator().manual_seed(rows + cols)
    x = torch.randn(rows, cols + 8, generator=g)
    got = ops.colsum(x.to(DEV)[:, :cols])            # a row-strided view, like the halves of a k | v gradient
    assert rel(got, x[:, :cols].double().sum(0)) < 1e-5


GEOMS = [  # (N, segment lengths) -- the reference's edge geometries: N around segment lengths,
    # N % r != 0, N < 16, tails that are almost all padding
    (7, [16, 24, 32, 64, 128]), (75, [16, 24, 32, 64, 128]), (97, [8, 32, 64, 96, 256]),
    (130, [32, 64, 128, 256, 512]), (1024, None), (1025, None), (1500, None), (2049, [256, 512, 1024, 4096, 8192]),
]


def _qkv(N, n_alloc, g, dt, scale=1.0):
    qkv = torch.zeros(n_alloc, 2304)
    qkv[:N] = torch.randn(N, 2304, generator=g) * scale
    return qkv.to(dt)


@pytest.mark.parametrize("N,sl", GEOMS)
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_dilated_attention_simt_vs_oracle(N, sl, dt):
    sl = sl or optimal_segment_lengths()
    geom = ops.Geometry.get(N, sl, DILATED_RATIO)
    g = torch.Generator().manual_seed(N)
    qkv = _qkv(N, geom.n_alloc, g, dt, 1.5)
    gamma, beta = 1 + 0.1 * torch.randn(768, generator=g), 0.1 * torch.randn(768, generator=g)
    dy = torch.randn(N, 768, generator=g).to(dt)
    o_c, l_c = C.dilated_attn_fwd(geom, qkv, 0)
    o, l = ops.dilated_attn_fwd(geom, qkv.to(DEV), 0)
    assert rel(l, l_c) < 1e-5 and rel(o, o_c) < tol(dt)
    y_c, a_c, lse_c, m_c, r_c = C.dilated_merge_ln_fwd(geom, o_c, l_c, gamma, beta, want_attn=True)
    y, a, lse, m, r = ops.dilated_merge_ln_fwd(geom, o, l, gamma.to(DEV), beta.to(DEV), want_attn=True)
    assert rel(a, a_c) < tol(dt) and rel(lse, lse_c) < 1e-5 and rel(y, y_c) < 2 * tol(dt)
    da_c, de_c = C.dilated_merge_ln_bwd(geom, dy, o_c, l_c, gamma, m_c, r_c)
    da, de = ops.dilated_merge_ln_bwd(geom, dy.to(DEV), o, l, gamma.to(DEV), m, r)
    assert da.shape[0] == geom.n_alloc and float(da[N:].abs().sum()) == 0.0
    assert rel(da, da_c) < 2 * tol(dt) and rel(de, de_c) < 2 * tol(dt)
    dq_c = C.dilated_attn_bwd(geom, qkv, da_c, lse_c, de_c, 0)
    dq = ops.dilated_attn_bwd(geom, qkv.to(DEV), da_c.to(DEV), lse_c.to(DEV), de_c.to(DEV), 0)
    assert rel(dq, dq_c) < (1e-4 if dt == torch.float32 else 3e-2)


def test_dilated_attention_linearity_in_v_at_full_size():
    """Size-independent property at the bench size (10k tiles): attention is linear in V, and the merged weights sum
    to one (constant V -> constant output)."""
    N = 10001
    geom = ops.Geometry.get(N, optimal_segment_lengths(), DILATED_RATIO)
    g = torch.Generator().manual_seed(1)
    qkv = _qkv(N, geom.n_alloc, g, torch.float32).to(DEV)
    ones = torch.ones(768, device=DEV)
    zeros = torch.zeros(768, device=DEV)

    def attn(t):
        o, l = ops.dilated_attn_fwd(geom, t, 0)
        return ops.dilated_merge_ln_fwd(geom, o, l, ones, zeros, want_attn=True)[1]

    a1 = attn(qkv)
    q2 = qkv.clone()
    q2[:N, 1536:] *= -2.5
    assert rel(attn(q2), -2.5 * a1) < 1e-5
    q3 = qkv.clone()
    q3[:N, 1536:] = 0.75
    # rows whose branches contain zero-padded slots see value 0 there: output <= 0.75, and == 0.75 where no padding
    a3 = attn(q3)
    assert float(a3.max()) <= 0.75 + 1e-5 and float(a3.min()) > 0.0


# ---------------------------------------------------------------------------------------------------------------------
# tcgen05 / TMA dilated attention (impl = 1) against the fp32-math SIMT kernels and the oracle
# ---------------------------------------------------------------------------------------------------------------------
# sizes at which the tcgen05 kernels are held against the ORACLE itself (not only against the SIMT kernels): every small
# geometry plus the bench geometries of BASELINE configs 2 and 3 (the oracle core needs 2 - 25 s of host time there)
ORACLE_SIZES = {n for n, _ in GEOMS} | {5793, 10001, 32769}
FWD_IMPLS = (1, 3)   # tcgen05 variants of mt_dilated_attn_fwd: 128-key and 48-key (default) score tiles (0 = SIMT cross-check)
SM100_GEOMS = GEOMS + [(5793, None), (10001, None), (300, [128, 256, 512, 1024, 2048]),
                       # whole-tile padding skip: last segments with 1 / 128 / 129 real slots, real counts that are
                       # exact multiples of the 128-slot tile, tails of several all-padding tiles in every branch
                       (1153, None), (2177, [1024, 2048, 4096, 8192, 16384]), (8193, None), (16385, None),
                       (4224, [512, 1024, 2048, 4096, 8192]), (12289, [4096, 4096, 8192, 8192, 16384]),
                       (32769, None),   # C3: 33 x 1024, 6 x 5792 and the 2 x 32768 branch whose tail segment is all padding
                       (40001, None)]   # C5 upper end


@pytest.mark.parametrize("N,sl", SM100_GEOMS)
def test_dilated_attention_tcgen05_forward(N, sl):
    sl = sl or optimal_segment_lengths()
    geom = ops.Geometry.get(N, sl, DILATED_RATIO)
    g = torch.Generator().manual_seed(N + 1)
    qkv = _qkv(N, geom.n_alloc, g, torch.bfloat16, 1.5).to(DEV)
    o_s, l_s = ops.dilated_attn_fwd(geom, qkv, 0)
    oc = None
    for impl in FWD_IMPLS:
        o_t, l_t = ops.dilated_attn_fwd(geom, qkv, impl)
        torch.cuda.synchronize()
        assert rel(l_t, l_s) < 2e-3, (impl, rel(l_t, l_s))       # P is rounded to bf16 before P V; lse itself is fp32
        assert float((l_t - l_s).abs().max()) < 2e-2, impl
        assert rel(o_t, o_s) < 3e-2, (impl, rel(o_t, o_s))
        if N in ORACLE_SIZES:
            if oc is None:
                oc = C.dilated_attn_fwd(geom, qkv.cpu(), 0)
            assert rel(l_t, oc[1]) < 2e-3 and rel(o_t, oc[0]) < 3e-2, impl


@pytest.mark.parametrize("N", [1025, 5793])
def test_dilated_attention_tcgen05_forward_rising_maximum(N):
    """Key magnitudes grow along the sequence, so the running row maximum jumps by far more than the lazy-rescale
    threshold (2^8) in later key tiles: the in-TMEM rescale of the O accumulator must kick in."""
    geom = ops.Geometry.get(N, optimal_segment_lengths(), DILATED_RATIO)
    g = torch.Generator().manual_seed(7 * N)
    qkv = torch.zeros(geom.n_alloc, 2304)
    qkv[:N] = torch.randn(N, 2304, generator=g)
    ramp = (1.0 + 8.0 * torch.arange(N) / N).unsqueeze(1)            # keys 9x larger at the end of the sequence
    qkv[:N, 768:1536] *= ramp
    sign = torch.where(torch.arange(N) % 256 < 128, 1.0, -1.0).unsqueeze(1)
    qkv[:N, 768:1536] *= sign                                         # and alternating, so maxima move both ways
    qkv = qkv.to(torch.bfloat16).to(DEV)
    o_s, l_s = ops.dilated_attn_fwd(geom, qkv, 0)
    for impl in FWD_IMPLS:
        o_t, l_t = ops.dilated_attn_fwd(geom, qkv, impl)
        torch.cuda.synchronize()
        assert torch.isfinite(o_t.float()).all() and torch.isfinite(l_t).all(), impl
        assert float((l_t - l_s).abs().max()) < 5e-2, (impl, float((l_t - l_s).abs().max()))
        assert rel(o_t, o_s) < 3e-2, (impl, rel(o_t, o_s))


BWD_IMPLS = (1, 2)      # one CTA per key tile, persistent CTAs with overlapped item hand-over


@pytest.mark.parametrize("N,sl", SM100_GEOMS)
def test_dilated_attention_tcgen05_backward(N, sl):
    sl = sl or optimal_segment_lengths()
    geom = ops.Geometry.get(N, sl, DILATED_RATIO)
    g = torch.Generator().manual_seed(N + 2)
    qkv = _qkv(N, geom.n_alloc, g, torch.bfloat16, 1.5).to(DEV)
    gamma, beta = (1 + 0.1 * torch.randn(768, generator=g)).to(DEV), (0.1 * torch.randn(768, generator=g)).to(DEV)
    dy = torch.randn(N, 768, generator=g).to(DEV)
    o, l = ops.dilated_attn_fwd(geom, qkv, 0)
    y, _, lse, m, r = ops.dilated_merge_ln_fwd(geom, o, l, gamma, beta)
    dattn, delta = ops.dilated_merge_ln_bwd(geom, dy, o, l, gamma, m, r)
    dq_s = ops.dilated_attn_bwd(geom, qkv, dattn, lse, delta, 0)
    dq_c = None
    if N in ORACLE_SIZES:   # the oracle's autograd through the dilated core with the same dO (cpu_kernels.dilated_attn_bwd)
        dq_c = C.dilated_attn_bwd(geom, qkv.cpu(), dattn.cpu(), lse.cpu(), delta.cpu(), 0)
    for impl in BWD_IMPLS:
        dq_t = ops.dilated_attn_bwd(geom, qkv, dattn, lse, delta, impl)
        torch.cuda.synchronize()
        for name, sl_ in (("dq", slice(0, 768)), ("dk", slice(768, 1536)), ("dv", slice(1536, 2304))):
            e = rel(dq_t[:, sl_], dq_s[:, sl_])
            assert e < 3e-2, (impl, name, e)
            if dq_c is not None:
                e = rel(dq_t[:, sl_], dq_c[:, sl_])
                assert e < 3e-2, (impl, name, "vs oracle", e)
                ct = torch.nn.functional.cosine_similarity(dq_t[:, sl_].flatten().double().cpu(),
                                                           dq_c[:, sl_].flatten().double(), dim=0)
                assert float(ct) > 0.9995, (impl, name, float(ct))


# ---------------------------------------------------------------------------------------------------------------------
# tcgen05 GEMM of the frozen linear layers with fused epilogues (mt_linear_sm100) against plain PyTorch fp32
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("impl", [1, 2])      # 1 = one CTA per 128 x 256 tile, 2 = CTA pairs (cta_group::2)
@pytest.mark.parametrize("M,N,K", [(10001, 2304, 768), (777, 768, 3072), (128, 256, 64), (1, 768, 768), (4097, 3072, 768),
                                   (257, 256, 128), (32769, 768, 768)])
def test_linear_sm100_plain_bias_residual(M, N, K, impl):
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).to(DEV)
    w = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(M, N, generator=g).to(DEV)
    ref = a.float() @ w.float().t()
    o32, o16 = ops.linear_sm100(a, w, want_f32=True, want_bf16=True, impl=impl)
    assert rel(o32, ref) < 1e-5 and rel(o16, ref) < 1e-2
    o32, _ = ops.linear_sm100(a, w, bias=bias, residual=res, impl=impl)
    assert rel(o32, ref + bias + res) < 1e-5
    # strided output: the QKV buffer has n_alloc rows, the GEMM writes the first M
    buf = torch.zeros(M + 7, N, device=DEV, dtype=torch.bfloat16)
    ops.linear_sm100(a, w, bias=bias, want_f32=False, out_bf16=buf[:M], impl=impl)
    assert rel(buf[:M], ref + bias) < 1e-2 and float(buf[M:].abs().sum()) == 0.0


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("M", [10001, 513])
def test_linear_sm100_ffn_epilogues(M, impl):
    """fc1 + GELU + LN statistics in one GEMM, fc2 with ffn_layernorm folded in + residual in the other: together they
    must equal fc2(LN(gelu(fc1(x)))) + residual of feedforward_network.py:132-143 (fp32 math on the same bf16 operands)."""
    g = torch.Generator().manual_seed(M)
    x = torch.randn(M, 768, generator=g).to(torch.bfloat16).to(DEV)
    w1 = (torch.randn(3072, 768, generator=g) * 0.04).to(DEV)
    b1 = (torch.randn(3072, generator=g) * 0.1).to(DEV)
    w2 = (torch.randn(768, 3072, generator=g) * 0.02).to(DEV)
    b2 = (torch.randn(768, generator=g) * 0.1).to(DEV)
    gam = (1 + 0.1 * torch.randn(3072, generator=g)).to(DEV)
    bet = (0.1 * torch.randn(3072, generator=g)).to(DEV)
    res = torch.randn(M, 768, generator=g).to(DEV)
    w1b = w1.to(torch.bfloat16)
    h_ref = x.float() @ w1b.float().t() + b1
    u_ref = torch.nn.functional.gelu(h_ref)
    stats = torch.full((M, 24, 2), float("nan"), device=DEV)     # every slab partial is written, nothing is accumulated
    h, u = ops.linear_sm100(x, w1b, mode=_lib.MT_EPI_GELU_STATS, bias=b1, want_f32=True, want_bf16=True, stats=stats, impl=impl)
    h2, u2 = ops.linear_sm100(x, w1b, mode=_lib.MT_EPI_GELU_STATS, bias=b1, want_f32=True, want_bf16=True, stats=stats.clone(), impl=impl)
    assert torch.equal(h, h2) and torch.equal(u, u2)             # bit-reproducible
    assert rel(h, h_ref) < 1e-5
    assert float((u.float() - u_ref).abs().max()) < 2e-2 * float(u_ref.abs().max())
    assert float((u.float() - u_ref.to(torch.bfloat16).float()).abs().max()) < 1e-2 * float(u_ref.abs().max())  # <= 1 bf16 ulp
    assert rel(stats[:, :, 0].sum(1), u.float().sum(1)) < 1e-4 and rel(stats[:, :, 1].sum(1), u.float().square().sum(1)) < 1e-4
    w2g = (w2 * gam[None, :]).to(torch.bfloat16)
    c1 = w2g.float().sum(1)
    c2 = w2 @ bet + b2
    y, _ = ops.linear_sm100(u, w2g, mode=_lib.MT_EPI_LN_RESIDUAL, residual=res, stats=stats, col_c1=c1, col_c2=c2,
                            ln_cols=3072, impl=impl)
    y2, _ = ops.linear_sm100(u, w2g, mode=_lib.MT_EPI_LN_RESIDUAL, residual=res, stats=stats, col_c1=c1, col_c2=c2,
                             ln_cols=3072, impl=impl)
    assert torch.equal(y, y2)
    ln = torch.nn.functional.layer_norm(u_ref, (3072,), gam, bet, 1e-5)
    y_ref = ln @ w2.t() + b2 + res
    assert rel(y, y_ref) < 1e-2        # bf16 operands (u, W2 gamma) against the fp32 chain
    # and against the same bf16 operands in fp32 math: the fold itself is exact up to fp32 rounding
    mean = u.float().mean(1, keepdim=True)
    rstd = torch.rsqrt(u.float().var(1, unbiased=False, keepdim=True) + 1e-5)
    y_same = rstd * (u.float() @ w2g.float().t() - mean * c1) + c2 + res
    assert rel(y, y_same) < 2e-4
