"""Parity of the sm_100a path (through the C ABI, kvpress_b200.native -> libkvpress_b200.so) against
the CPU oracle and the golden vectors of the imported reference. Needs a B200: `pytest -m gpu`.

Bars (north_star): Knorm / StreamingLLM — identical retained-index sets (tie-aware where the
reference's own top-k is ambiguous, see oracle.check_selection); attention-based scorers — scores
within 1e-3 relative of the fp32 evaluation of the reference formula, and within 16-bit rounding
noise (<= 4 ulp, 99.9% <= 2 ulp) of the reference's own 16-bit scores.
"""
import pytest
import torch

from oracle import press_oracle as O
from tests.conftest import ulp16_diff

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _native():
    from kvpress_b200 import native
    native.load()
    return native


def _check_compaction(k, v, k_out, v_out, idx):
    """K'/V' must be exactly the rows the returned indices name, ascending, unique."""
    idx = idx.long().cpu()
    assert (idx[..., 1:] > idx[..., :-1]).all(), "indices must be strictly ascending"
    assert torch.equal(k_out.cpu(), O.gather_rows(k.cpu(), idx))
    assert torch.equal(v_out.cpu(), O.gather_rows(v.cpu(), idx))


# ---------------------------------------------------------------------------------------------------
# Knorm
# ---------------------------------------------------------------------------------------------------
def test_knorm_scores_vs_golden(golden):
    nat = _native()
    k = golden.t("keys").to(DEV)
    got = nat.knorm_score(k).cpu()
    ref = golden.t("knorm_scores")
    d = ulp16_diff(got, ref)
    # fp32 summation order may flip a 16-bit rounding on ~1e-5 of elements, never by more than 1 ulp
    assert d.max() <= 1
    assert (d > 0).float().mean() < 1e-3


def test_knorm_compress_vs_golden(golden):
    nat = _native()
    k, v = golden.t("keys").to(DEV), golden.t("values").to(DEV)
    for i, r in enumerate(golden.ratios):
        n_kept = O.kept_count(golden.S, r)
        k_out, v_out, idx, scores = nat.knorm_compress(k, v, n_kept, return_indices=True, return_scores=True)
        assert k_out.shape == (golden.B, golden.Hkv, n_kept, golden.D)
        _check_compaction(k, v, k_out, v_out, idx)
        # the kept set is exactly the canonical (lowest-position ties) selection of the kernel's own scores
        assert torch.equal(idx.cpu().long(), O.select_lowest_index_ties(scores.cpu(), n_kept))
        # and it is a valid top-k of the REFERENCE's scores
        res = O.check_selection(golden.t("knorm_scores"), idx.cpu(), n_kept, ulp_slack=1)
        assert res["ok"], res
        # identical to the reference's own retained set wherever the reference had no ties to break
        ref_sc = golden.t("knorm_scores")
        if torch.equal(scores.cpu(), ref_sc):
            canon = O.select_lowest_index_ties(ref_sc, n_kept)
            assert torch.equal(idx.cpu().long(), canon)


@pytest.mark.parametrize("shape", [(1, 8, 32768, 128), (2, 4, 5000, 128), (3, 2, 1023, 64), (1, 1, 1025, 256),
                                   (1, 2, 777, 96), (2, 2, 64, 32), (1, 3, 1, 128)])
@pytest.mark.parametrize("ratio", [0.1, 0.5, 0.875])
def test_knorm_compress_vs_oracle_random(shape, ratio):
    nat = _native()
    torch.manual_seed(hash((shape, ratio)) % 2**31)
    k = torch.randn(shape, dtype=torch.bfloat16)
    v = torch.randn(shape, dtype=torch.bfloat16)
    n_kept = O.kept_count(shape[2], ratio)
    kd, vd = k.to(DEV), v.to(DEV)
    k_out, v_out, idx, scores = nat.knorm_compress(kd, vd, n_kept, return_indices=True, return_scores=True)
    assert k_out.shape[2] == n_kept
    if n_kept == 0:
        return
    _check_compaction(k, v, k_out, v_out, idx)
    ref_scores = O.knorm_scores(k)
    assert ulp16_diff(scores.cpu(), ref_scores).max() <= 1
    assert torch.equal(idx.cpu().long(), O.select_lowest_index_ties(scores.cpu(), n_kept))
    assert O.check_selection(ref_scores, idx.cpu(), n_kept, ulp_slack=1)["ok"]


def test_knorm_full_size_properties():
    """BASELINE shape [1,8,131072,128]: size-independent properties instead of an element-wise oracle —
    highest-score-kept invariant (reference tests/presses/test_presses.py:143-162), exact gather,
    idempotence of compaction at the same n_kept."""
    nat = _native()
    torch.manual_seed(7)
    S, n_kept = 131072, O.kept_count(131072, 0.5)
    k = torch.randn(1, 8, S, 128, dtype=torch.bfloat16, device=DEV)
    v = torch.randn(1, 8, S, 128, dtype=torch.bfloat16, device=DEV)
    k_out, v_out, idx, scores = nat.knorm_compress(k, v, n_kept, return_indices=True, return_scores=True)
    idxl = idx.long()
    assert (idxl[..., 1:] > idxl[..., :-1]).all()
    assert torch.equal(k_out, k.gather(2, idxl.unsqueeze(-1).expand(-1, -1, -1, 128)))
    assert torch.equal(v_out, v.gather(2, idxl.unsqueeze(-1).expand(-1, -1, -1, 128)))
    # kept scores are the n_kept largest: sorted kept scores == top of the sorted scores
    kept_scores = scores.float().gather(2, idxl).sort(-1).values
    top_scores = scores.float().sort(-1).values[..., -n_kept:]
    assert torch.equal(kept_scores, top_scores)
    # score tensor agrees with torch's own norm on the GPU to 1 ulp
    assert ulp16_diff(scores.cpu(), (-k.norm(dim=-1)).cpu()).max() <= 1
    # compacting the compacted cache with ratio 0 -> same rows (idempotence at n_kept == S')
    k2, v2, idx2, _ = nat.knorm_compress(k_out, v_out, n_kept, return_indices=True)
    assert torch.equal(k2, k_out) and torch.equal(v2, v_out)
    assert torch.equal(idx2.long(), torch.arange(n_kept, device=DEV).expand_as(idx2))


def test_knorm_strided_views_and_fp16():
    """Views as left by pipeline._remove_answer_from_cache (reference pipeline.py:252-265)."""
    nat = _native()
    torch.manual_seed(3)
    big_k = torch.randn(2, 4, 700, 128, dtype=torch.float16, device=DEV)
    big_v = torch.randn(2, 4, 700, 128, dtype=torch.float16, device=DEV)
    k, v = big_k[:, :, :611], big_v[:, 1:3, 5:616]
    k = k[:, 1:3]
    assert not k.is_contiguous()
    n_kept = 300
    k_out, v_out, idx, scores = nat.knorm_compress(k, v, n_kept, return_indices=True, return_scores=True)
    _check_compaction(k, v, k_out, v_out, idx)
    assert ulp16_diff(scores.cpu(), O.knorm_scores(k.cpu())).max() <= 1
    assert k_out.is_contiguous() and v_out.is_contiguous()


# ---------------------------------------------------------------------------------------------------
# StreamingLLM
# ---------------------------------------------------------------------------------------------------
def test_streaming_vs_golden(golden):
    nat = _native()
    k, v = golden.t("keys").to(DEV), golden.t("values").to(DEV)
    for i, r in enumerate(golden.ratios):
        n_kept = O.kept_count(golden.S, r)
        sc = nat.streaming_score(k, n_kept, 4)
        assert torch.equal(sc.cpu(), golden.t(f"streaming_scores_{i}"))
        k_out, v_out, idx = nat.streaming_compress(k, v, n_kept, 4, return_indices=True)
        assert torch.equal(idx.cpu(), golden.t(f"streaming_kept_{i}"))  # identical retained-index sets
        _check_compaction(k, v, k_out, v_out, idx)


@pytest.mark.parametrize("S,ratio,n_sink", [(131072, 0.5, 4), (1000, 0.999, 4), (10, 0.5, 4), (5000, 0.25, 0),
                                            (2049, 0.3, 128)])
def test_streaming_edge_cases(S, ratio, n_sink):
    nat = _native()
    k = torch.randn(1, 2, S, 128, dtype=torch.bfloat16, device=DEV)
    v = torch.randn(1, 2, S, 128, dtype=torch.bfloat16, device=DEV)
    n_kept = O.kept_count(S, ratio)
    k_out, v_out, idx = nat.streaming_compress(k, v, n_kept, n_sink, return_indices=True)
    if n_kept == 0:
        assert k_out.shape[2] == 0
        return
    want = O.streaming_kept(S, n_kept, n_sink).to(torch.int32)
    assert torch.equal(idx.cpu(), want.expand_as(idx))
    # equals top-k of the reference's 0/1 scores with the lowest-position tie rule
    sc = O.streaming_scores(k.cpu(), ratio, n_sink)
    assert torch.equal(O.select_lowest_index_ties(sc, n_kept), want.long().expand(1, 2, -1))
    _check_compaction(k, v, k_out, v_out, idx)


# ---------------------------------------------------------------------------------------------------
# generic scores -> select + compact (what wrapper presses / custom ScorerPress subclasses use)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["random", "all_equal", "two_values", "with_inf_nan_zero", "ascending"])
def test_scores_compress_tie_rule(kind):
    nat = _native()
    torch.manual_seed(5)
    B, H, S, D = 2, 3, 4099, 64
    k = torch.randn(B, H, S, D, dtype=torch.bfloat16)
    v = torch.randn(B, H, S, D, dtype=torch.bfloat16)
    if kind == "random":
        sc = torch.randn(B, H, S)
    elif kind == "all_equal":
        sc = torch.full((B, H, S), 0.5)
    elif kind == "two_values":
        sc = (torch.rand(B, H, S) > 0.5).float()
    elif kind == "ascending":
        sc = torch.arange(S).float().expand(B, H, S).clone()
    else:
        sc = torch.randn(B, H, S)
        sc[..., 10] = float("inf")
        sc[..., 11] = float("-inf")
        sc[..., 12] = float("nan")
        sc[..., 13:40] = 0.0
        sc[..., 40:60] = -0.0
    sc = sc.to(torch.bfloat16)
    for n_kept in (1, 7, S // 2, S - 1, S):
        k_out, v_out, idx = nat.scores_compress(sc.to(DEV), k.to(DEV), v.to(DEV), n_kept, return_indices=True)
        _check_compaction(k, v, k_out, v_out, idx)
        assert torch.equal(idx.cpu().long(), O.select_lowest_index_ties(sc, n_kept)), (kind, n_kept)


def test_errors_are_loud():
    nat = _native()
    k = torch.randn(1, 2, 128, 64, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        nat.knorm_compress(k, k, 64)
    kf = torch.randn(1, 2, 128, 64, dtype=torch.float32, device=DEV)
    with pytest.raises(RuntimeError, match="bf16/fp16"):
        nat.knorm_compress(kf, kf, 64)
    kd = torch.randn(1, 2, 128, 12, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(RuntimeError, match="unsupported shape"):
        nat.knorm_compress(kd, kd, 64)
