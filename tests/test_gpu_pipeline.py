"""End-to-end on a B200: the hook path (`with press(model)`), the "kv-press-text-generation" pipeline and
DecodingPress running through the CUDA library on random-init bf16 models (configs[0]/[3] in small)."""
import logging

import pytest
import torch
from transformers import DynamicCache

from kvpress_b200 import (DecodingPress, ExpectedAttentionPress, KnormPress, KVPressTextGenerationPipeline,
                          SnapKVPress, StreamingLLMPress, native)
from kvpress_b200.presses.scorer_press import kept_count
from oracle import press_oracle as O
from tests.conftest import ulp16_diff
from tests.tiny_models import tiny_llama, tiny_qwen3, word_tokenizer, words

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def model():
    return tiny_llama(dtype=torch.bfloat16, device=DEV, head_dim=64, heads=4, kv_heads=2)


@pytest.fixture(scope="module")
def qwen():
    return tiny_qwen3(dtype=torch.bfloat16, device=DEV, head_dim=128, heads=4, kv_heads=2)


def _full_cache(model, ids):
    cache = DynamicCache()
    model.model(input_ids=ids, past_key_values=cache)
    return cache


@pytest.mark.parametrize("ratio", [0.2, 0.5, 0.8])
def test_knorm_hook_keeps_highest_scores(model, ratio):
    """Reference tests/presses/test_presses.py:143-162 on the GPU path."""
    ids = torch.randint(2, 250, (3, 700), device=DEV)
    full = _full_cache(model, ids)
    cache = DynamicCache()
    with KnormPress(compression_ratio=ratio)(model):
        model.model(input_ids=ids, past_key_values=cache)
    n_kept = kept_count(700, ratio)
    assert cache.get_seq_length() == n_kept
    for lf, lc in zip(full.layers, cache.layers):
        scores = native.knorm_score(lf.keys)
        idx = O.select_lowest_index_ties(scores.cpu(), n_kept).to(DEV)
        gi = idx.unsqueeze(-1).expand(-1, -1, -1, lf.keys.shape[-1])
        assert torch.equal(lc.keys, lf.keys.gather(2, gi)) and torch.equal(lc.values, lf.values.gather(2, gi))
        assert lc.keys.is_contiguous()


@pytest.mark.parametrize("press", [StreamingLLMPress(0.5), SnapKVPress(0.5), ExpectedAttentionPress(0.7),
                                   ExpectedAttentionPress(0.3, use_covariance=False, use_vnorm=False)])
def test_hook_lengths_and_rows_are_cache_rows(model, press):
    ids = torch.randint(2, 250, (2, 900), device=DEV)
    full = _full_cache(model, ids)
    cache = DynamicCache()
    with press(model):
        model.model(input_ids=ids, past_key_values=cache)
    n_kept = kept_count(900, press.compression_ratio)
    assert cache.get_seq_length() == n_kept
    for lf, lc in zip(full.layers, cache.layers):
        assert lc.keys.shape == (2, 2, n_kept, 64)
        # every kept row is a row of the uncompressed cache, K and V from the same position, ascending
        sig_full = lf.keys.float().sum(-1) + 7 * lf.values.float().sum(-1)
        sig_kept = lc.keys.float().sum(-1) + 7 * lc.values.float().sum(-1)
        pos = torch.searchsorted(sig_full.sort(-1).values, sig_kept)
        assert (pos < 900).all()
        assert torch.isin(sig_kept[0, 0], sig_full[0, 0]).all()


class _Spy:
    def __init__(self):
        self.calls = []


def test_snapkv_and_ea_scores_through_hook_vs_oracle(model):
    """score() of the attention-based presses as wrappers call it: prologue on the GPU (torch), scan in the
    CUDA library; compared with the oracle fed the SAME prologue outputs."""
    ids = torch.randint(2, 250, (1, 1100), device=DEV)
    captured = {}

    class Grab(KnormPress):
        def compress(self, module, hidden_states, keys, values, attentions, kwargs):
            if module.layer_idx == 1:
                captured.update(module=module, hidden=hidden_states, keys=keys, values=values, kwargs=kwargs)
            return keys, values

    with Grab(compression_ratio=0.5)(model):
        model.model(input_ids=ids, past_key_values=DynamicCache())
    mod, hid, k, v, kw = (captured[n] for n in ("module", "hidden", "keys", "values", "kwargs"))

    snap = SnapKVPress(0.5)
    q_win = snap.window_queries(mod, hid, kw)
    got = snap.score(mod, hid, k, v, None, kw).cpu()
    hi = O.snapkv_scores_fp32(q_win.cpu(), k.cpu(), 64, 5)
    keep = slice(0, 1100 - 64)
    assert ulp16_diff(got[..., keep], hi[..., keep].to(torch.bfloat16)).max() <= 1

    ea = ExpectedAttentionPress(0.5)
    mu, cov = ea.get_query_statistics(mod, hid)
    got = ea.score(mod, hid, k, v, None, kw).cpu()
    hi = O.expected_attention_scores_fp32(k.cpu(), v.cpu(), mu.cpu(), cov.cpu(), 0.0, 4, True)
    assert ulp16_diff(got[..., 4:], hi[..., 4:].to(torch.bfloat16)).max() <= 1
    # host prologue == the oracle's restatement of the reference prologue (bf16 GEMM on GPU vs CPU)
    cos, sin = kw["position_embeddings"]
    q_ref = O.snapkv_window_queries(hid.cpu(), mod.q_proj.weight.detach().cpu(), 4, 64, cos.cpu(), sin.cpu(), 64)
    assert torch.allclose(q_win.cpu().float(), q_ref.float(), atol=0.06, rtol=0.03)


def test_pipeline_on_gpu(model, caplog):
    pipe = KVPressTextGenerationPipeline(model=model, tokenizer=word_tokenizer(), device=DEV)
    context = words(1000, seed=1)
    with caplog.at_level(logging.DEBUG):
        out = pipe(context, question=words(5, seed=2), press=KnormPress(compression_ratio=0.5), max_new_tokens=8)
    assert isinstance(out["answer"], str)
    msgs = [r.message for r in caplog.records]
    assert "Context Length: 1001" in msgs and "Compressed Context Length: 500" in msgs
    for press in (SnapKVPress(0.5), ExpectedAttentionPress(0.7), StreamingLLMPress(0.25)):
        out = pipe(context, questions=[words(3, seed=3), words(4, seed=4)], press=press, max_new_tokens=4)
        assert len(out["answers"]) == 2
    # ratio 0 == no press
    a = pipe(context, question=words(5, seed=2), press=KnormPress(0.0), max_new_tokens=8)["answer"]
    b = pipe(context, question=words(5, seed=2), max_new_tokens=8)["answer"]
    assert a == b


@pytest.mark.parametrize("base", [KnormPress, StreamingLLMPress, ExpectedAttentionPress])
def test_decoding_press_on_gpu(qwen, base):
    """configs[3] in small: DecodingPress(base, interval, target) on a Qwen3-style model (q_norm/k_norm)."""
    pipe = KVPressTextGenerationPipeline(model=qwen, tokenizer=word_tokenizer(), device=DEV)
    press = DecodingPress(base_press=base(), compression_interval=16, target_size=256)
    sizes = []
    orig = press.forward_hook

    def spy(module, inp, kwargs, output):
        out = orig(module, inp, kwargs, output)
        if module.layer_idx == 1:
            sizes.append(kwargs["past_key_values"].get_seq_length(1))
        return out

    press.forward_hook = spy
    pipe(words(400, seed=5), question=words(6, seed=6), press=press, max_new_tokens=80)
    assert len(sizes) >= 40
    assert min(sizes[16:]) >= 256 and max(sizes[16:]) <= 256 + 16 - 1 and 256 in sizes


@pytest.mark.parametrize("inner", [KnormPress(0.5), SnapKVPress(0.4), StreamingLLMPress(0.6)])
def test_key_rerotation_press_through_hook_vs_oracle(model, inner, monkeypatch):
    """KeyRerotationPress around a scorer press, through the forward hook on a bf16 model: V rows are cache rows,
    K rows are the oracle's re-rotation of the kept cache rows."""
    from kvpress_b200 import KeyRerotationPress, native

    kept, real = [], native.scores_compress_rerotate

    def spy(scores, keys, values, n_kept, inv_freq, return_indices=False):
        k, v, idx = real(scores, keys, values, n_kept, inv_freq, return_indices=True)
        kept.append(idx.long().cpu())
        return k, v, idx

    monkeypatch.setattr(native, "scores_compress_rerotate", spy)
    ids = torch.randint(2, 250, (2, 700), device=DEV)
    full = _full_cache(model, ids)
    cache = DynamicCache()
    with KeyRerotationPress(inner)(model):
        model.model(input_ids=ids, past_key_values=cache)
    n_kept = kept_count(700, inner.compression_ratio)
    assert cache.get_seq_length() == n_kept and len(kept) == len(cache.layers)
    inv_freq = model.model.rotary_emb.inv_freq.float().cpu()
    for lf, lc, pos in zip(full.layers, cache.layers, kept):
        assert (pos[..., 1:] > pos[..., :-1]).all()
        assert torch.equal(O.gather_rows(lf.values.cpu(), pos), lc.values.cpu())
        ref_k = O.rerotate_keys(lf.keys.cpu(), pos, inv_freq)
        exact = (ref_k.view(torch.int16) == lc.keys.cpu().view(torch.int16)).float().mean().item()
        assert exact > 0.99, exact
        assert (ref_k.float() - lc.keys.cpu().float()).abs().max().item() <= 2.0 ** -6 * lf.keys.float().abs().max().item()


@pytest.mark.parametrize("n_tokens,chunk", [(900, 256), (1024, 256), (700, 1024)])
def test_chunk_press_on_gpu_matches_per_chunk_oracle(model, n_tokens, chunk):
    """ChunkPress(KnormPress): all full chunks go through ONE kvp_scores_compress call on the strided
    [B*H, n_chunks, L, D] view of the cache; the result must be the per-chunk oracle selection, concatenated."""
    from kvpress_b200 import ChunkPress

    ids = torch.randint(2, 250, (2, n_tokens), device=DEV)
    full = _full_cache(model, ids)
    cache = DynamicCache()
    with ChunkPress(KnormPress(0.5), chunk_length=chunk)(model):
        model.model(input_ids=ids, past_key_values=cache)
    for lf, lc in zip(full.layers, cache.layers):
        K, V = lf.keys.cpu(), lf.values.cpu()
        ks, vs = [], []
        for lo in range(0, n_tokens, chunk):
            kc, vc = K[:, :, lo:lo + chunk], V[:, :, lo:lo + chunk]
            n_kept = max(1, int(kc.shape[2] * 0.5))
            idx = O.select_lowest_index_ties(O.knorm_scores(kc), n_kept)
            ks.append(O.gather_rows(kc, idx))
            vs.append(O.gather_rows(vc, idx))
        assert torch.equal(lc.keys.cpu(), torch.cat(ks, dim=2)) and torch.equal(lc.values.cpu(), torch.cat(vs, dim=2))


def test_pyramidkv_and_wrappers_on_gpu(model):
    from kvpress_b200 import ComposedPress, PerLayerCompressionPress, PyramidKVPress, RandomPress
    from kvpress_b200.presses.pyramidkv_press import pyramid_layer_budget

    ids = torch.randint(2, 250, (2, 800), device=DEV)
    n_layers = model.config.num_hidden_layers
    cache = DynamicCache()
    with PyramidKVPress(0.6, window_size=32, beta=4)(model):
        model.model(input_ids=ids, past_key_values=cache)
    assert [la.keys.shape[2] for la in cache.layers] == [pyramid_layer_budget(800, 0.6, 32, 4, n_layers, i)
                                                         for i in range(n_layers)]
    cache = DynamicCache()
    with PerLayerCompressionPress(KnormPress(), [0.25, 0.75][:n_layers] + [0.5] * (n_layers - 2))(model):
        model.model(input_ids=ids, past_key_values=cache)
    assert cache.layers[0].keys.shape[2] == 600 and cache.layers[1].keys.shape[2] == 200
    cache = DynamicCache()
    press = ComposedPress([RandomPress(0.5, seed=1), KnormPress(0.5)])
    with press(model):
        model.model(input_ids=ids, past_key_values=cache)
    assert cache.get_seq_length() == 200 and press.compression_ratio == pytest.approx(0.75)
