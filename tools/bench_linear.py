"""GPU: the hand-written tcgen05 GEMM (mt_linear_sm100) against cuBLAS on the frozen-linear shapes of one encoder layer."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from modaltune_b200 import _lib, ops  # noqa: E402

dev = "cuda"
M = int(sys.argv[1]) if len(sys.argv) > 1 else 10001


def med(fn, reps=9):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


g = torch.Generator().manual_seed(0)
for name, N, K in (("qkv", 2304, 768), ("out_proj", 768, 768), ("fc1", 3072, 768), ("fc2", 768, 3072), ("dX qkv", 768, 2304)):
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).to(dev)
    w = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    o32 = torch.empty(M, N, device=dev)
    o16 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    fl = 2.0 * M * N * K
    t_cublas32 = med(lambda: torch.mm(a, w.t(), out_dtype=torch.float32))
    t_cublas16 = med(lambda: torch.addmm(bias.to(torch.bfloat16), a, w.t()))
    print(f"{name:9s} M={M} N={N} K={K}: cuBLAS f32-out {t_cublas32*1e3:7.1f} us ({fl/t_cublas32/1e9:6.0f} TF)  bf16+bias {t_cublas16*1e3:7.1f} us", end="")
    for impl in (1, 2):
        t_ours32 = med(lambda: ops.linear_sm100(a, w, out_f32=o32, impl=impl))
        t_ours16 = med(lambda: ops.linear_sm100(a, w, bias=bias, want_f32=False, out_bf16=o16, impl=impl))
        print(f" | ours[{'1cta' if impl == 1 else 'pair'}] f32-out {t_ours32*1e3:7.1f} us ({fl/t_ours32/1e9:6.0f} TF)  bf16+bias {t_ours16*1e3:7.1f} us ({fl/t_ours16/1e9:6.0f} TF)", end="")
    print()
# the FFN pair with fused epilogues against the current path (GEMM + gelu_ln kernel + GEMM + residual kernel)
x = torch.randn(M, 768, generator=g).to(torch.bfloat16).to(dev)
w1 = (torch.randn(3072, 768, generator=g) * 0.04).to(torch.bfloat16).to(dev)
w2 = (torch.randn(768, 3072, generator=g) * 0.02).to(torch.bfloat16).to(dev)
b1, b2 = torch.randn(3072, generator=g).to(dev), torch.randn(768, generator=g).to(dev)
gam, bet = torch.ones(3072, device=dev), torch.zeros(3072, device=dev)
c1, c2 = w2.float().sum(1), b2.clone()
res = torch.randn(M, 768, generator=g).to(dev)
stats = torch.empty(M, 24, 2, device=dev)


def ffn_old():
    f1 = torch.mm(x, w1.t(), out_dtype=torch.float32)
    gg, _, _ = ops.gelu_ln_fwd(f1, gam, bet, torch.bfloat16, hbias=b1)
    f2 = torch.mm(gg, w2.t(), out_dtype=torch.float32)
    return ops.residual_bias_add(res, f2, b2)


def ffn_new(impl):
    h, u = ops.linear_sm100(x, w1, mode=_lib.MT_EPI_GELU_STATS, bias=b1, want_f32=True, want_bf16=True, stats=stats, impl=impl)
    return ops.linear_sm100(u, w2, mode=_lib.MT_EPI_LN_RESIDUAL, residual=res, stats=stats, col_c1=c1, col_c2=c2, ln_cols=3072, impl=impl)[0]


t_old = med(ffn_old)
for impl in (1, 2):
    t_new = med(lambda: ffn_new(impl))
    err = float((ffn_old() - ffn_new(impl)).abs().max() / ffn_old().abs().max())
    print(f"FFN forward (fc1, GELU, LN 3072, fc2, residual) M={M}: cuBLAS + element-wise kernels {t_old*1e3:.1f} us, fused tcgen05 GEMMs "
          f"[{'1cta' if impl == 1 else 'pair'}] {t_new*1e3:.1f} us, speed-up {t_old/t_new:.2f}x, max rel diff {err:.2e}")
