"""Length-aware sharding of variable-length slides over data-parallel ranks (BASELINE config 5).

The reference shards cases with a ``DistributedSampler`` (``utils/base_trainer.py:283-286``): every rank gets the same
NUMBER of slides.  With 2k-40k tiles per slide the cost of a step varies by > 20x, so the slowest rank sets the pace of
every all-reduce.  Here the slides of one global step are assigned by estimated cost instead (longest-processing-time
first onto the least loaded rank); ranks may then hold different numbers of slides and all-reduce once per global step.

The cost model follows the measured structure of a step: a part linear in the token count (GEMMs, LayerNorm / GELU /
residual kernels, adapters) and the dilated-attention part, proportional to the algorithmic attention FLOPs of the slide's
geometry.  The two rates are the round-1 measurements at 10k tiles; only their ratio matters for the assignment.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

HEADS, HEAD_DIM = 16, 48
DILATED_RATIO = (1, 2, 4, 8, 16)
SEGMENT_LENGTHS = (1024, 5792, 32768, 185363, 1048576)   # optimal_segment_lengths(262144, 256), slide_encoder.py

# ms per token of the linear part and ms per algorithmic attention GFLOP (forward + backward = 3.5 x forward FLOPs),
# from the 10k-tile step of DESIGN.md section 6: 59.5 ms, of which 21.4 ms attention at 12.03 TFLOP per step
MS_PER_TOKEN = (59.5 - 21.4) / 10001.0
MS_PER_ATTN_GFLOP = 21.4 / 12030.0


def attention_gflop(n_tokens: int, segment_lengths: Sequence[int] = SEGMENT_LENGTHS,
                    ratios: Sequence[int] = DILATED_RATIO) -> float:
    """Algorithmic attention GFLOP of one slide step (3 task passes x 12 layers x 3.5 x forward): 4 d sum c^2 over
    (branch, segment, head) with c = real positions the head owns."""
    fwd = 0.0
    for sl, r in zip(segment_lengths, ratios):
        g = min(int(sl), n_tokens)
        n_seg = -(-n_tokens // g)
        for s in range(n_seg):
            lo, hi = s * g, min(n_tokens, (s + 1) * g)
            for h in range(HEADS):
                off = (h * r) // HEADS
                c = max(0, -(-(hi - lo - off) // r))
                fwd += 4.0 * HEAD_DIM * c * c
    return fwd * 3.5 * 12 * 3 / 1e9


def estimate_step_ms(n_tiles: int) -> float:
    """Estimated forward + backward time of one slide with ``n_tiles`` tiles on one B200 (bf16 mode)."""
    n = n_tiles + 1
    return MS_PER_TOKEN * n + MS_PER_ATTN_GFLOP * attention_gflop(n)


def pack_slides(tile_counts: Sequence[int], world_size: int) -> Tuple[List[List[int]], Dict[str, float]]:
    """Assign slide indices to ``world_size`` ranks, balancing the estimated step time (LPT greedy: slides by decreasing
    cost, each onto the currently least loaded rank; ties keep rank order so the assignment is deterministic on every
    rank without communication).  Returns (per-rank index lists, {"makespan_ms", "mean_ms", "imbalance"})."""
    assert world_size >= 1
    costs = [estimate_step_ms(int(t)) for t in tile_counts]
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += costs[i]
    mean = sum(loads) / world_size
    return shards, {"makespan_ms": max(loads) if loads else 0.0, "mean_ms": mean,
                    "imbalance": (max(loads) / mean - 1.0) if mean > 0 else 0.0}


def round_robin(tile_counts: Sequence[int], world_size: int) -> Tuple[List[List[int]], Dict[str, float]]:
    """The reference's equal-count sharding (DistributedSampler without shuffling), for comparison."""
    shards = [list(range(r, len(tile_counts), world_size)) for r in range(world_size)]
    loads = [sum(estimate_step_ms(int(tile_counts[i])) for i in sh) for sh in shards]
    mean = sum(loads) / world_size
    return shards, {"makespan_ms": max(loads), "mean_ms": mean, "imbalance": max(loads) / mean - 1.0 if mean > 0 else 0.0}
