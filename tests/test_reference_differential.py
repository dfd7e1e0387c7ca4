"""Differential test against the UNMODIFIED reference, imported from /root/reference (build container
only; skipped on the GPU box where it does not exist): the re-written presses, driven through the same
hooks on the same random-init model, must retain the same rows as the reference presses."""
import os
import sys
import types

import pytest
import torch
from transformers import DynamicCache

from kvpress_b200 import ExpectedAttentionPress, KeyDiffPress, KnormPress, SnapKVPress, StreamingLLMPress
from tests import cpu_backend
from tests.tiny_models import tiny_llama, tiny_qwen3

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    sys.modules.setdefault("fire", types.ModuleType("fire"))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import kvpress

    return kvpress


def _rows_signature(keys, values):
    """Order-independent signature of the retained (K, V) rows of every head."""
    g = torch.Generator().manual_seed(0)
    w = torch.randn(keys.shape[-1], generator=g, dtype=torch.float64)
    sig = keys.double() @ w + 3.0 * (values.double() @ w)
    return sig.sort(-1).values


CASES = [
    ("knorm", lambda m: m.KnormPress, KnormPress, {}),
    ("streaming", lambda m: m.StreamingLLMPress, StreamingLLMPress, {"n_sink": 4}),
    ("snapkv", lambda m: m.SnapKVPress, SnapKVPress, {"window_size": 16, "kernel_size": 5}),
    ("expected_attention", lambda m: m.ExpectedAttentionPress, ExpectedAttentionPress, {"n_sink": 4}),
    ("keydiff", lambda m: m.KeyDiffPress, KeyDiffPress, {}),
    ("tova", lambda m: m.TOVAPress, __import__("kvpress_b200").TOVAPress, {}),
]


@pytest.mark.parametrize("name,ref_cls,our_cls,kw", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("ratio", [0.25, 0.5, 0.7])
@pytest.mark.parametrize("family", ["llama", "qwen3"])
def test_same_rows_as_reference(monkeypatch, ref, name, ref_cls, our_cls, kw, ratio, family):
    cpu_backend.install(monkeypatch)
    model = tiny_llama() if family == "llama" else tiny_qwen3()
    # distinct tokens per sequence: a repeated token gives layer-0 keys of exactly equal norm (RoPE keeps
    # norms), i.e. exact score ties that torch.topk and the kernels may legitimately break differently
    g = torch.Generator().manual_seed(5)
    ids = torch.stack([torch.randperm(250, generator=g)[:160] + 2 for _ in range(2)])
    S = ids.shape[1]

    theirs = DynamicCache()
    with ref_cls(ref)(compression_ratio=ratio, **kw)(model):
        # transformers >= 5.5 no longer hands cache_position to attention; the reference hook needs it
        model.model(input_ids=ids, past_key_values=theirs, cache_position=torch.arange(S))
    ours = DynamicCache()
    with our_cls(compression_ratio=ratio, **kw)(model):
        model.model(input_ids=ids, past_key_values=ours)

    assert ours.get_seq_length() == theirs.get_seq_length() == int(S * (1 - ratio))
    for lo, lt in zip(ours.layers, theirs.layers):
        assert lo.keys.shape == lt.keys.shape
        if name == "knorm":
            # Qwen3's k_norm makes every key norm (nearly) equal: exact ties, broken differently by
            # torch.topk and by the lowest-position rule. The multiset of kept scores must still agree.
            assert torch.allclose(lo.keys.norm(dim=-1).sort(-1).values, lt.keys.norm(dim=-1).sort(-1).values)
        else:
            assert torch.allclose(_rows_signature(lo.keys, lo.values), _rows_signature(lt.keys, lt.values),
                                  atol=1e-9)


@pytest.mark.parametrize("inner", ["streaming", "snapkv"])
@pytest.mark.parametrize("family", ["llama", "qwen3"])
def test_key_rerotation_same_cache_as_reference(monkeypatch, ref, inner, family):
    """KeyRerotationPress (key_rerotation_press.py:133-152): the reference sorts the kept indices, so the
    compacted cache must agree ROW BY ROW, re-rotated keys included."""
    from kvpress_b200 import KeyRerotationPress

    cpu_backend.install(monkeypatch)
    model = tiny_llama() if family == "llama" else tiny_qwen3()
    g = torch.Generator().manual_seed(7)
    ids = torch.stack([torch.randperm(250, generator=g)[:160] + 2 for _ in range(2)])
    S, ratio = ids.shape[1], 0.4
    _, ref_cls, our_cls, kw = next(c for c in CASES if c[0] == inner)

    theirs = DynamicCache()
    with ref.KeyRerotationPress(ref_cls(ref)(compression_ratio=ratio, **kw))(model):
        model.model(input_ids=ids, past_key_values=theirs, cache_position=torch.arange(S))
    ours = DynamicCache()
    press = KeyRerotationPress(our_cls(compression_ratio=ratio, **kw))
    assert press.compression_ratio == ratio
    with press(model):
        model.model(input_ids=ids, past_key_values=ours)

    assert ours.get_seq_length() == theirs.get_seq_length() == int(S * (1 - ratio))
    for lo, lt in zip(ours.layers, theirs.layers):
        assert torch.equal(lo.values, lt.values)
        assert torch.equal(lo.keys, lt.keys)


def _prefill_both(ref_press, our_press, model, ids):
    S = ids.shape[1]
    theirs, ours = DynamicCache(), DynamicCache()
    with ref_press(model):
        model.model(input_ids=ids, past_key_values=theirs, cache_position=torch.arange(S))
    with our_press(model):
        model.model(input_ids=ids, past_key_values=ours)
    return theirs, ours


def _distinct_ids(seed, n=160, batch=2):
    g = torch.Generator().manual_seed(seed)
    return torch.stack([torch.randperm(250, generator=g)[:n] + 2 for _ in range(batch)])


def _assert_same_rows(ours, theirs):
    for lo, lt in zip(ours.layers, theirs.layers):
        assert lo.keys.shape == lt.keys.shape
        assert torch.allclose(_rows_signature(lo.keys, lo.values), _rows_signature(lt.keys, lt.values), atol=1e-9)


@pytest.mark.parametrize("ratio,beta", [(0.3, 20), (0.6, 4), (0.5, 1), (0.9, 20)])
def test_pyramidkv_same_rows_and_budgets_as_reference(monkeypatch, ref, ratio, beta):
    """pyramidkv_press.py:47-112: per-layer budgets (incl. the fallback branch) and the kept rows."""
    from kvpress_b200 import PyramidKVPress
    from kvpress_b200.presses.pyramidkv_press import pyramid_layer_budget

    cpu_backend.install(monkeypatch)
    model = tiny_llama(layers=4)
    ids = _distinct_ids(11)
    kw = dict(compression_ratio=ratio, window_size=16, kernel_size=5, beta=beta)
    theirs, ours = _prefill_both(ref.PyramidKVPress(**kw), PyramidKVPress(**kw), model, ids)
    lens = [layer.keys.shape[2] for layer in ours.layers]
    assert lens == [layer.keys.shape[2] for layer in theirs.layers]
    n_layers = model.config.num_hidden_layers
    assert lens == [pyramid_layer_budget(160, ratio, 16, beta, n_layers, i) for i in range(n_layers)]
    _assert_same_rows(ours, theirs)
    # the budget formula alone, over a grid, against the reference method
    from types import SimpleNamespace
    for q_len in (100, 777, 4096, 131072):
        for r in (0.1, 0.5, 0.75, 0.95):
            for b in (1, 5, 20):
                for layer in (0, 7, 31):
                    mod = SimpleNamespace(config=SimpleNamespace(num_hidden_layers=32), layer_idx=layer)
                    want = ref.PyramidKVPress(compression_ratio=r, window_size=64, beta=b).get_layer_budget(mod, q_len)
                    assert pyramid_layer_budget(q_len, r, 64, b, 32, layer) == want


@pytest.mark.parametrize("inner", ["knorm", "snapkv"])
@pytest.mark.parametrize("n_tokens,chunk", [(160, 40), (150, 64), (100, 256)])
def test_chunk_press_same_rows_as_reference(monkeypatch, ref, inner, n_tokens, chunk):
    from kvpress_b200 import ChunkPress

    cpu_backend.install(monkeypatch)
    model = tiny_llama()
    ids = _distinct_ids(12, n=n_tokens)
    _, ref_cls, our_cls, kw = next(c for c in CASES if c[0] == inner)
    if inner == "snapkv":
        kw = {"window_size": 8, "kernel_size": 3}
    theirs, ours = _prefill_both(ref.ChunkPress(ref_cls(ref)(compression_ratio=0.5, **kw), chunk_length=chunk),
                                 ChunkPress(our_cls(compression_ratio=0.5, **kw), chunk_length=chunk), model, ids)
    assert ours.get_seq_length() == theirs.get_seq_length()
    _assert_same_rows(ours, theirs)


def test_composed_and_per_layer_wrappers_same_rows_as_reference(monkeypatch, ref):
    from kvpress_b200 import ComposedPress, PerLayerCompressionPress

    cpu_backend.install(monkeypatch)
    model = tiny_llama()
    ids = _distinct_ids(13)
    # second press = Knorm: its scores do not depend on the ROW ORDER the first press leaves behind (the reference
    # emits top-k order, this package ascending positions; index-based scorers such as SnapKV's window or
    # StreamingLLM's sinks would read a different "last w rows" from the reference's permuted cache)
    rp = ref.ComposedPress([ref.SnapKVPress(0.25, window_size=16), ref.KnormPress(0.4)])
    op = ComposedPress([SnapKVPress(0.25, window_size=16), KnormPress(0.4)])
    theirs, ours = _prefill_both(rp, op, model, ids)
    assert ours.get_seq_length() == theirs.get_seq_length()
    assert op.compression_ratio == pytest.approx(rp.compression_ratio)
    _assert_same_rows(ours, theirs)

    ratios = [0.2, 0.6][: model.config.num_hidden_layers] + [0.5] * max(0, model.config.num_hidden_layers - 2)
    rp = ref.PerLayerCompressionPress(ref.SnapKVPress(window_size=16), ratios)
    op = PerLayerCompressionPress(SnapKVPress(window_size=16), ratios)
    theirs, ours = _prefill_both(rp, op, model, ids)
    assert [la.keys.shape[2] for la in ours.layers] == [la.keys.shape[2] for la in theirs.layers]
    assert op.compression_ratio == pytest.approx(rp.compression_ratio)
    with pytest.raises(AttributeError):
        op.compression_ratio = 0.1
    _assert_same_rows(ours, theirs)


def test_random_press_same_rows_as_reference_on_cpu(monkeypatch, ref):
    from kvpress_b200 import RandomPress

    cpu_backend.install(monkeypatch)
    model = tiny_llama()
    ids = _distinct_ids(14)
    theirs, ours = _prefill_both(ref.RandomPress(0.5, seed=3), RandomPress(0.5, seed=3), model, ids)
    _assert_same_rows(ours, theirs)


@pytest.mark.parametrize("inner", ["expected_attention", "snapkv"])
@pytest.mark.parametrize("alpha", [0.2, 0.0, 1.0])
def test_adakv_masks_the_same_keys_as_reference(monkeypatch, ref, inner, alpha):
    """adakv_press.py:53-78: per-head safeguard + bottom-k across heads; the pruned (batch, head, position)
    triples must be the reference's, and one decoding step through attention_patch must give the same output."""
    from kvpress_b200 import AdaKVPress

    cpu_backend.install(monkeypatch)
    model = tiny_llama()
    ids = _distinct_ids(21)
    S = ids.shape[1]
    _, ref_cls, our_cls, kw = next(c for c in CASES if c[0] == inner)
    attns = [layer.self_attn for layer in model.model.layers]

    def masks():
        out = []
        for a in attns:
            b, h, s = a.masked_key_indices
            out.append(set(zip(b.tolist(), h.tolist(), s.tolist())))
        return out

    def run(press, with_cache_position):
        cache = DynamicCache()
        with press(model):
            extra = {"cache_position": torch.arange(S)} if with_cache_position else {}
            model.model(input_ids=ids, past_key_values=cache, **extra)
            m = masks()
            extra = {"cache_position": torch.tensor([S])} if with_cache_position else {}
            step = model.model(input_ids=ids[:, :1], past_key_values=cache, **extra).last_hidden_state
        for a in attns:
            a.masked_key_indices = None
        return cache, m, step

    c_ref, m_ref, y_ref = run(ref.AdaKVPress(ref_cls(ref)(compression_ratio=0.5, **kw), alpha_safeguard=alpha), True)
    c_our, m_our, y_our = run(AdaKVPress(our_cls(compression_ratio=0.5, **kw), alpha_safeguard=alpha), False)
    H = model.config.num_key_value_heads
    for a, b in zip(m_our, m_ref):
        assert len(a) == len(b) == ids.shape[0] * H * (S - int(S * 0.5))
        assert a == b
    assert c_our.get_seq_length() == c_ref.get_seq_length() == S + 1      # AdaKV never shrinks the cache
    assert torch.allclose(y_our, y_ref, atol=1e-5)


@pytest.mark.parametrize("inner", ["knorm", "snapkv", "expected_attention"])
@pytest.mark.parametrize("n_tokens,chunk", [(160, 20), (150, 20), (150, 32), (15, 20)])
def test_chunkkv_press_same_rows_as_reference(monkeypatch, ref, inner, n_tokens, chunk):
    """chunkkv_press.py:52-117: whole chunks kept, ranked by head-summed mean score; ragged tail; S < chunk."""
    from kvpress_b200 import ChunkKVPress

    if inner != "knorm" and n_tokens < 32:
        pytest.skip("window scorers need more tokens than their window")
    cpu_backend.install(monkeypatch)
    model = tiny_llama()
    ids = _distinct_ids(31, n=n_tokens)
    _, ref_cls, our_cls, kw = next(c for c in CASES if c[0] == inner)
    if inner == "snapkv":
        kw = {"window_size": 8, "kernel_size": 3}
    theirs, ours = _prefill_both(ref.ChunkKVPress(ref_cls(ref)(compression_ratio=0.4, **kw), chunk_length=chunk),
                                 ChunkKVPress(our_cls(compression_ratio=0.4, **kw), chunk_length=chunk), model, ids)
    assert ours.get_seq_length() == theirs.get_seq_length()
    if n_tokens < chunk:                               # plain press.compress: same rows, reference in top-k order
        _assert_same_rows(ours, theirs)
        return
    for lo, lt in zip(ours.layers, theirs.layers):     # both emit position order: caches are equal row by row
        assert torch.equal(lo.keys, lt.keys) and torch.equal(lo.values, lt.values)


@pytest.mark.parametrize("inner", ["knorm", "keydiff"])
@pytest.mark.parametrize("block_size", [16, 50, 400])
def test_block_press_same_rows_as_reference(monkeypatch, ref, inner, block_size):
    """block_press.py:49-98 incl. the reference's own invariant (tests/presses/test_block_press.py:30-63):
    a block at least as long as the context is the plain press."""
    from kvpress_b200 import BlockPress

    cpu_backend.install(monkeypatch)
    model = tiny_llama()
    ids = _distinct_ids(32)
    _, ref_cls, our_cls, kw = next(c for c in CASES if c[0] == inner)
    theirs, ours = _prefill_both(ref.BlockPress(ref_cls(ref)(compression_ratio=0.5, **kw), block_size=block_size),
                                 BlockPress(our_cls(compression_ratio=0.5, **kw), block_size=block_size), model, ids)
    _assert_same_rows(ours, theirs)
    if block_size >= ids.shape[1]:
        plain = DynamicCache()
        with our_cls(compression_ratio=0.5, **kw)(model):
            model.model(input_ids=ids, past_key_values=plain)
        _assert_same_rows(ours, plain)


def test_expected_attention_stats_press_same_rows_and_folder_layout_as_reference(monkeypatch, ref, tmp_path):
    """expected_attention_with_stats.py:21-110: stored (mu, Sigma) per layer instead of the per-prompt prologue;
    statistics folders are interchangeable with the reference's PyTorchModelHubMixin layout; the offline collector
    reproduces the reference formula (mean, unbiased covariance of the pre-RoPE queries beyond n_sink)."""
    from kvpress_b200 import ExpectedAttentionStatsPress
    from kvpress_b200.presses.expected_attention_with_stats import ExpectedAttentionStats, collect_query_statistics
    from kvpress.presses.expected_attention_with_stats import ExpectedAttentionStats as RefStats

    cpu_backend.install(monkeypatch)
    model = tiny_llama()
    cfg = model.config
    L, Hq, D = cfg.num_hidden_layers, cfg.num_attention_heads, cfg.head_dim
    batches = [_distinct_ids(40 + i, n=120, batch=1) for i in range(3)]
    stats = collect_query_statistics(model, batches, n_sink=4)
    # independent restatement of the reference's collect_queries arithmetic
    from kvpress_b200.utils import get_prerope_query_states
    qs = [[] for _ in range(L)]
    hooks = [layer.self_attn.register_forward_pre_hook(
        lambda m, a, kw, i=i: qs[i].append(get_prerope_query_states(m, kw["hidden_states"])[:, :, 4:]), with_kwargs=True)
        for i, layer in enumerate(model.model.layers)]
    with torch.no_grad():
        for ids in batches:
            model.model(input_ids=ids)
    for h in hooks:
        h.remove()
    for i in range(L):
        q = torch.cat(qs[i], dim=-2)[0]                                     # [Hq, N, D]
        mean = q.mean(dim=-2)
        c = q - mean.unsqueeze(-2)
        assert torch.allclose(stats.query_mean[i], mean, atol=1e-5)
        assert torch.allclose(stats.query_cov[i], c.transpose(-2, -1) @ c / (q.shape[-2] - 1), atol=1e-5)

    # folder layout: ours -> reference loader, reference -> our loader
    stats.save_pretrained(str(tmp_path / "ours"))
    theirs = RefStats.from_pretrained(str(tmp_path / "ours"))
    assert torch.equal(theirs.query_mean.data, stats.query_mean.data) and torch.equal(theirs.query_cov.data, stats.query_cov.data)
    theirs.save_pretrained(str(tmp_path / "theirs"))
    back = ExpectedAttentionStats.from_pretrained(str(tmp_path / "theirs"))
    assert torch.equal(back.query_cov.data, stats.query_cov.data) and back.meta["n_sink"] == 4

    # the presses: same rows, with and without covariance; hidden states are not needed
    ids = _distinct_ids(50)
    for use_cov in (True, False):
        rp = ref.ExpectedAttentionStatsPress(compression_ratio=0.5, use_covariance=use_cov, stats_folder=str(tmp_path / "ours"))
        op = ExpectedAttentionStatsPress(compression_ratio=0.5, use_covariance=use_cov, stats_folder=str(tmp_path / "theirs"))
        t_cache, o_cache = _prefill_both(rp, op, model, ids)
        assert torch.equal(op.mu, rp.mu) and op.mu.shape == (L, Hq, D)
        _assert_same_rows(o_cache, t_cache)
    assert ExpectedAttentionStatsPress.needs_hidden_states is False
    with pytest.raises(ValueError, match="No statistics given"):
        ExpectedAttentionStatsPress(0.5).post_init_from_model(model)


# ---------------------------------------------------------------------------------------------------------------------
# pipeline-level differentials: the reference pipeline driven on the same model, with a pre-hook that
# restores the `cache_position` kwarg transformers >= 5.3 no longer hands to attention modules
# ---------------------------------------------------------------------------------------------------------------------
def _restore_cache_position(model):
    handles = []

    def pre(module, args, kwargs):
        if kwargs.get("cache_position") is None:
            past = kwargs["past_key_values"].get_seq_length(module.layer_idx)
            kwargs["cache_position"] = torch.arange(past, past + kwargs["hidden_states"].shape[1])
        return args, kwargs

    for layer in model.model.layers:
        handles.append(layer.self_attn.register_forward_pre_hook(pre, with_kwargs=True))
    return handles


def _both_pipelines(ref, model):
    from kvpress_b200 import KVPressTextGenerationPipeline
    from tests.tiny_models import word_tokenizer

    tok = word_tokenizer()
    return ref.KVPressTextGenerationPipeline(model=model, tokenizer=tok), KVPressTextGenerationPipeline(model=model, tokenizer=tok)


@pytest.mark.parametrize("family", ["llama", "qwen3"])
def test_pipeline_with_key_rerotation_matches_reference(monkeypatch, ref, family):
    """pipeline.py:231-232 of the reference: after KeyRerotationPress the question / answer positions continue from
    the COMPRESSED length. Same answers, and the positions the model sees are the reference's."""
    from kvpress_b200 import KeyRerotationPress
    from tests.tiny_models import words

    cpu_backend.install(monkeypatch)
    model = tiny_llama() if family == "llama" else tiny_qwen3()
    ref_pipe, our_pipe = _both_pipelines(ref, model)
    context, questions = words(150, seed=3), [words(4, seed=4), words(6, seed=5)]
    seen = {"ref": [], "ours": []}

    def spy(tag):
        def pre(module, args, kwargs):
            if kwargs.get("position_ids") is not None:
                seen[tag].append(kwargs["position_ids"].flatten().tolist())
        return model.model.register_forward_pre_hook(pre, with_kwargs=True)

    handles = _restore_cache_position(model) + [spy("ref")]
    try:
        theirs = ref_pipe(context, questions=questions, max_new_tokens=6,
                          press=ref.KeyRerotationPress(ref.StreamingLLMPress(compression_ratio=0.4, n_sink=4)))
    finally:
        for h in handles:
            h.remove()
    h = spy("ours")
    try:
        ours = our_pipe(context, questions=questions, max_new_tokens=6,
                        press=KeyRerotationPress(StreamingLLMPress(compression_ratio=0.4, n_sink=4)))
    finally:
        h.remove()
    assert ours["answers"] == theirs["answers"]
    assert seen["ours"] == seen["ref"] and len(seen["ours"]) > 2
    n_kept = int(151 * (1 - 0.4))
    assert seen["ours"][0][0] == n_kept          # the first question token sits right after the compacted cache


def test_decoding_press_with_stats_press_matches_reference(monkeypatch, ref, tmp_path):
    """DecodingPress(ExpectedAttentionStatsPress): the stats press reads only the LENGTH of the buffered hidden states
    (future RoPE positions); the zero-copy length-only buffer must reproduce the reference's concatenated buffer."""
    from kvpress_b200 import DecodingPress, ExpectedAttentionStatsPress
    from kvpress_b200.presses.expected_attention_with_stats import collect_query_statistics
    from tests.tiny_models import words

    cpu_backend.install(monkeypatch)
    model = tiny_llama()
    stats = collect_query_statistics(model, [_distinct_ids(60 + i, n=100, batch=1) for i in range(2)], n_sink=4)
    stats.save_pretrained(str(tmp_path / "stats"))
    ref_pipe, our_pipe = _both_pipelines(ref, model)
    context, question = words(60, seed=8), words(5, seed=9)
    kw = dict(compression_interval=7, target_size=40, hidden_states_buffer_size=256)
    # few future positions: the average RoPE rotation then depends visibly on where they start (q_len)
    skw = dict(stats_folder=str(tmp_path / "stats"), n_future_positions=3)

    rp = ref.DecodingPress(base_press=ref.ExpectedAttentionStatsPress(**skw), **kw)
    theirs_cache, ours_cache = DynamicCache(), DynamicCache()
    handles = _restore_cache_position(model)
    try:
        theirs = ref_pipe(context, question=question, max_new_tokens=24, press=rp, cache=theirs_cache)
    finally:
        for h in handles:
            h.remove()
    op = DecodingPress(base_press=ExpectedAttentionStatsPress(**skw), **kw)
    ours = our_pipe(context, question=question, max_new_tokens=24, press=op, cache=ours_cache)
    assert ours["answer"] == theirs["answer"]
    assert [la.keys.shape[2] for la in ours_cache.layers] == [la.keys.shape[2] for la in theirs_cache.layers]
    _assert_same_rows(ours_cache, theirs_cache)
    assert not op.hidden_states_buffer and not op.hidden_states_lens     # reset on exit, nothing was cloned
