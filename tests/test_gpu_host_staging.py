"""kvpress_b200.host_staging.compress_host (pinned host K,V -> pinned host K',V', three-stream pipeline) against
the device-resident C-ABI calls on the same data."""
import pytest
import torch

from oracle import press_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _inputs(B, H, Hq, S, D, seed=0):
    g = torch.Generator().manual_seed(seed)
    K = torch.randn(B, H, S, D, generator=g).to(torch.bfloat16).pin_memory()
    V = torch.randn(B, H, S, D, generator=g).to(torch.bfloat16).pin_memory()
    q = torch.randn(B, Hq, 16, D, generator=g).to(torch.bfloat16).to(DEV)
    mu = (0.5 * torch.randn(B, Hq, D, generator=g)).to(torch.bfloat16).to(DEV)
    a = torch.randn(B, Hq, D, D, generator=g) / D ** 0.5
    cov = (a @ a.transpose(-1, -2)).to(torch.bfloat16).to(DEV)
    return K, V, q, mu, cov


@pytest.mark.parametrize("heads_per_chunk", [1, 3])
@pytest.mark.parametrize("zero_copy", [False, True])
def test_knorm_and_streaming_host_path_equal_device_path(heads_per_chunk, zero_copy):
    from kvpress_b200 import host_staging, native
    B, H, S, D = 2, 4, 5000, 128
    K, V, *_ = _inputs(B, H, H, S, D)
    n_kept = O.kept_count(S, 0.6)
    Kd, Vd = K.to(DEV), V.to(DEV)
    for scorer, direct in (("knorm", lambda: native.knorm_compress(Kd, Vd, n_kept)[:2]),
                           ("streaming", lambda: native.streaming_compress(Kd, Vd, n_kept, 4)[:2])):
        k_ref, v_ref = direct()
        k2, v2 = host_staging.compress_host(scorer, K, V, n_kept, device=DEV, heads_per_chunk=heads_per_chunk,
                                            values_zero_copy=zero_copy, n_sink=4)
        assert k2.is_pinned() and v2.is_pinned()
        assert torch.equal(k2, k_ref.cpu()) and torch.equal(v2, v_ref.cpu()), scorer
    # and against the oracle's selection rule
    idx = O.select_lowest_index_ties(O.knorm_scores(K), n_kept)
    k2, v2 = host_staging.compress_host("knorm", K, V, n_kept, device=DEV, values_zero_copy=zero_copy)
    assert torch.equal(k2, O.gather_rows(K, idx)) and torch.equal(v2, O.gather_rows(V, idx))


@pytest.mark.parametrize("zero_copy", [False, True])
def test_snapkv_host_path_equals_per_head_device_calls(zero_copy):
    from kvpress_b200 import host_staging, native
    B, H, Hq, S, D = 1, 4, 16, 3000, 128
    K, V, q, _, _ = _inputs(B, H, Hq, S, D, seed=1)
    n_kept = O.kept_count(S, 0.5)
    k2, v2 = host_staging.compress_host("snapkv", K, V, n_kept, device=DEV, values_zero_copy=zero_copy,
                                        q_window=q, window=16, kernel_size=5)
    G = Hq // H
    for h in range(H):
        kr, vr = native.snapkv_compress(K[:, h:h + 1].to(DEV), V[:, h:h + 1].to(DEV), q[:, h * G:(h + 1) * G], 16, 5,
                                        n_kept)[:2]
        assert torch.equal(k2[:, h:h + 1], kr.cpu()) and torch.equal(v2[:, h:h + 1], vr.cpu())


def test_expected_attention_host_path_equals_per_head_device_calls():
    from kvpress_b200 import host_staging, native
    B, H, Hq, S, D = 2, 2, 8, 2500, 128
    K, V, _, mu, cov = _inputs(B, H, Hq, S, D, seed=2)
    n_kept = O.kept_count(S, 0.7)
    k2, v2 = host_staging.compress_host("expected_attention", K, V, n_kept, device=DEV, mu=mu, cov=cov, epsilon=0.0,
                                        n_sink=4, use_vnorm=True)
    G = Hq // H
    for b in range(B):
        for h in range(H):
            kr, vr = native.expected_attention_compress(
                K[b:b + 1, h:h + 1].to(DEV), V[b:b + 1, h:h + 1].to(DEV), mu[b:b + 1, h * G:(h + 1) * G],
                cov[b:b + 1, h * G:(h + 1) * G], 0.0, 4, True, n_kept)[:2]
            assert torch.equal(k2[b:b + 1, h:h + 1], kr.cpu()) and torch.equal(v2[b:b + 1, h:h + 1], vr.cpu())
    with pytest.raises(RuntimeError):
        host_staging.compress_host("expected_attention", K, V, n_kept, device=DEV, mu=mu, cov=cov, values_zero_copy=True)


def test_host_path_rejects_unpinned_and_device_inputs():
    from kvpress_b200 import host_staging
    K = torch.randn(1, 1, 64, 64).to(torch.bfloat16)
    with pytest.raises(RuntimeError):
        host_staging.compress_host("knorm", K, K, 32, device=DEV)
    with pytest.raises(RuntimeError):
        host_staging.compress_host("knorm", K.to(DEV), K.to(DEV), 32, device=DEV)
