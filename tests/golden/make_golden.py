"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference)
on CPU for seeded inputs. Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Harness shims: a stub `fire` module (kvpress imports it at package import for an
unrelated CLI) — nothing in the reference is modified. 16-bit tensors are stored as uint16 views.
"""
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = "/root/reference"
OUT = Path(__file__).resolve().parent


def import_reference():
    sys.modules.setdefault("fire", types.ModuleType("fire"))
    sys.path.insert(0, REF)
    import kvpress  # noqa

    return kvpress


def u16(t: torch.Tensor) -> np.ndarray:
    assert t.dtype in (torch.bfloat16, torch.float16)
    return t.contiguous().view(torch.uint16).numpy()


CASES = {
    "small64": dict(B=2, Hq=4, Hkv=2, D=64, hidden=256, S=384, seed=11),
    "llama128": dict(B=1, Hq=8, Hkv=2, D=128, hidden=512, S=1200, seed=12),
    "ties128": dict(B=1, Hq=4, Hkv=1, D=128, hidden=256, S=2500, seed=13, heavy_tail=True),
    "half64": dict(B=1, Hq=2, Hkv=2, D=64, hidden=128, S=300, seed=14, dtype=torch.float16),
}


def make_case(name, **kw):
    out = build_case(**kw)
    np.savez_compressed(OUT / f"{name}.npz", **out)
    print(name, {k: v.shape for k, v in out.items() if k.endswith("scores")})


def build_case(*, B, Hq, Hkv, D, hidden, S, seed, dtype=torch.bfloat16, heavy_tail=False):
    """Runs the imported reference on one seeded case and returns the arrays of its .npz (also used LIVE by
    tests/test_oracle_golden.py: big bf16 GEMMs round differently on different host CPUs, so the stored files pin
    the oracle bit-exactly only on a host whose GEMM kernels match the generating one)."""
    kvpress = import_reference()
    from transformers import LlamaConfig
    from transformers.models.llama.modeling_llama import LlamaAttention, LlamaRotaryEmbedding

    torch.manual_seed(seed)
    cfg = LlamaConfig(hidden_size=hidden, num_attention_heads=Hq, num_key_value_heads=Hkv, head_dim=D,
                      num_hidden_layers=1, intermediate_size=2 * hidden, vocab_size=128,
                      max_position_embeddings=8192)
    cfg._attn_implementation = "sdpa"
    attn = LlamaAttention(cfg, 0).to(dtype).eval()
    attn.rotary_emb = LlamaRotaryEmbedding(cfg)
    hidden_states = torch.randn(B, S, hidden).to(dtype)
    keys = torch.randn(B, Hkv, S, D).to(dtype)
    values = torch.randn(B, Hkv, S, D).to(dtype)
    if heavy_tail:  # a few huge-norm "sink" keys and exact duplicates -> forced ties
        keys[:, :, :3] *= 10
        keys[:, :, 50:60] = keys[:, :, 40:50]
    cos, sin = attn.rotary_emb(hidden_states, torch.arange(S)[None])
    pe = (cos, sin)
    kwargs = {"position_embeddings": pe}
    out = {
        "meta": np.array([B, Hq, Hkv, D, hidden, S, seed], dtype=np.int64),
        "hidden_states": u16(hidden_states), "keys": u16(keys), "values": u16(values),
        "q_weight": u16(attn.q_proj.weight.detach()), "cos": u16(cos), "sin": u16(sin),
    }
    ratios = [0.1, 0.25, 0.5, 0.7, 0.875]
    out["ratios"] = np.array(ratios)

    with torch.no_grad():
        # ---- Knorm ---------------------------------------------------------------------------
        press = kvpress.KnormPress()
        out["knorm_scores"] = u16(press.score(attn, hidden_states, keys, values, None, kwargs))
        # ---- StreamingLLM ----------------------------------------------------------------------
        for i, r in enumerate(ratios):
            press = kvpress.StreamingLLMPress(compression_ratio=r, n_sink=4)
            sc = press.score(attn, hidden_states, keys, values, None, kwargs)
            out[f"streaming_scores_{i}"] = u16(sc)
        # ---- SnapKV ------------------------------------------------------------------------------
        from kvpress.presses.snapkv_press import SnapKVPress
        from kvpress.utils import get_prerope_query_states
        from transformers.models.llama.modeling_llama import rotate_half

        w = 64 if S > 128 else 16
        snap = SnapKVPress(window_size=w, kernel_size=5)
        q = get_prerope_query_states(attn, hidden_states[:, -w:])
        q_window = (q * cos[:, -w:].unsqueeze(1)) + (rotate_half(q) * sin[:, -w:].unsqueeze(1))
        out["snap_window"] = np.array([w, 5])
        out["snap_q_window"] = u16(q_window)
        out["snap_scores"] = u16(snap.score(attn, hidden_states, keys, values, None, kwargs))
        # ---- ExpectedAttention ---------------------------------------------------------------------
        ea = kvpress.ExpectedAttentionPress(n_sink=4, n_future_positions=512, use_covariance=True, use_vnorm=True)
        mu, cov = ea.get_query_statistics(attn, hidden_states)
        out["ea_mu"], out["ea_cov"] = u16(mu), u16(cov)
        fut = torch.arange(S, S + 512)[None]
        cf, sf = attn.rotary_emb(mu, fut)
        out["ea_cos_future"], out["ea_sin_future"] = u16(cf[0]), u16(sf[0])
        out["ea_scores"] = u16(ea.score(attn, hidden_states, keys, values, None, kwargs))
        ea2 = kvpress.ExpectedAttentionPress(n_sink=4, use_covariance=False, use_vnorm=False, epsilon=0.0)
        out["ea_scores_nocov_novnorm"] = u16(ea2.score(attn, hidden_states, keys, values, None, kwargs))
        ea3 = kvpress.ExpectedAttentionPress(n_sink=4, use_covariance=True, use_vnorm=True, epsilon=1e-2)
        out["ea_scores_eps"] = u16(ea3.score(attn, hidden_states, keys, values, None, kwargs))

        # ---- full compress(): kept index sets (sorted) and gathered rows checksum -----------------
        for i, r in enumerate(ratios):
            for tag, press in (
                ("knorm", kvpress.KnormPress(compression_ratio=r)),
                ("streaming", kvpress.StreamingLLMPress(compression_ratio=r, n_sink=4)),
                ("snap", SnapKVPress(compression_ratio=r, window_size=w, kernel_size=5)),
                ("ea", kvpress.ExpectedAttentionPress(compression_ratio=r)),
            ):
                sc = press.score(attn, hidden_states, keys, values, None, kwargs)
                n_kept = int(S * (1 - r))
                idx = sc.topk(n_kept, dim=-1).indices
                k2, v2 = press.compress(attn, hidden_states, keys, values, None, kwargs)
                assert k2.shape[2] == n_kept
                # the reference's own output must be the gather of its own top-k
                g = keys.gather(2, idx.unsqueeze(-1).expand(-1, -1, -1, D))
                assert torch.equal(g, k2)
                out[f"{tag}_kept_{i}"] = idx.sort(-1).values.numpy().astype(np.int32)
    return out


def make_decoding_table():
    """decoding_press.py:194-236 — ratio bisection table."""
    kvpress = import_reference()
    press = kvpress.DecodingPress(base_press=kvpress.KnormPress(), compression_interval=4, target_size=54)
    rows = []
    for q_len, target in [(58, 54), (108, 54), (2560, 2048), (4607, 2048), (4096 + 37 + 511, 2048), (100, 200),
                          (131072, 39321), (2049, 2048), (3000, 1), (7, 3)]:
        r = press._find_target_compression_ratio(q_len, target)
        rows.append((q_len, target, r, int(q_len * (1 - r))))
    np.savez(OUT / "decoding_ratio.npz", table=np.array(rows, dtype=np.float64))
    print("decoding table", rows[:3])


def make_rerotation():
    """key_rerotation_press.py:133-152 run unmodified. Two wrapped scorers per case: the
    reference KnormPress (realistic, ties in bf16) and a ScorerPress subclass whose scores are a seeded
    permutation of DISTINCT 16-bit values, so the kept set is unambiguous and outputs compare row by row."""
    kvpress = import_reference()
    from dataclasses import dataclass
    from types import SimpleNamespace

    @dataclass
    class PermutationPress(kvpress.ScorerPress):
        scores: torch.Tensor = None

        def score(self, module, hidden_states, keys, values, attentions, kwargs):
            return self.scores

    out = {}
    cases = [("a", 2, 2, 384, 64, torch.bfloat16, 10000.0), ("b", 1, 2, 700, 128, torch.bfloat16, 500000.0),
             ("c", 1, 2, 300, 64, torch.float16, 10000.0)]
    out["cases"] = np.array([[B, H, S, D, int(dt == torch.float16)] for _, B, H, S, D, dt, _ in cases], dtype=np.int64)
    out["ratios"] = np.array([0.25, 0.5, 0.8])
    for tag, B, H, S, D, dtype, theta in cases:
        torch.manual_seed(100 + S)
        keys, values = torch.randn(B, H, S, D).to(dtype), torch.randn(B, H, S, D).to(dtype)
        inv_freq = 1.0 / (theta ** (torch.arange(0, D, 2, dtype=torch.int64).float() / D))
        module = SimpleNamespace(rotary_emb=SimpleNamespace(inv_freq=inv_freq), head_dim=D)
        perm = torch.stack([torch.randperm(S) for _ in range(B * H)]).view(B, H, S)
        scores = (perm + 0x3000).to(torch.int16).view(dtype)  # distinct positive finite values
        assert scores.float().flatten(0, 1).unique(dim=-1).shape[-1] == S
        out[f"{tag}_keys"], out[f"{tag}_values"] = u16(keys), u16(values)
        out[f"{tag}_inv_freq"], out[f"{tag}_scores"] = inv_freq.numpy(), u16(scores)
        for i, r in enumerate(out["ratios"]):
            r = float(r)
            with torch.no_grad():
                press = kvpress.KeyRerotationPress(PermutationPress(compression_ratio=r, scores=scores))
                k2, v2 = press.compress(module, None, keys, values, None, {})
                out[f"{tag}_perm_k_{i}"] = u16(k2)
                if i == 0:
                    out[f"{tag}_perm_v_{i}"] = u16(v2)
                kn = kvpress.KnormPress(compression_ratio=r)
                sc = kn.score(module, None, keys, values, None, {})
                idx = sc.topk(int(S * (1 - r)), dim=-1).indices.sort(dim=2).values
                k3, v3 = kvpress.KeyRerotationPress(kn).compress(module, None, keys, values, None, {})
                out[f"{tag}_knorm_idx_{i}"] = idx.numpy().astype(np.int32)
                assert torch.equal(v3, values.gather(2, idx.unsqueeze(-1).expand(-1, -1, -1, D)))
                out[f"{tag}_knorm_k_{i}"] = u16(k3)
    np.savez_compressed(OUT / "rerotation.npz", **out)
    print("rerotation", out["cases"].tolist())


def make_keydiff():
    """keydiff_press.py:36-46 run unmodified: scores for three seeded caches (incl. a zero key and duplicates)."""
    kvpress = import_reference()
    out = {}
    cases = [("a", 2, 2, 384, 64, torch.bfloat16), ("b", 1, 2, 1000, 128, torch.bfloat16), ("c", 1, 2, 300, 64, torch.float16)]
    out["cases"] = np.array([[B, H, S, D, int(dt == torch.float16)] for _, B, H, S, D, dt in cases], dtype=np.int64)
    for tag, B, H, S, D, dtype in cases:
        torch.manual_seed(200 + S)
        keys = (torch.randn(B, H, S, D) + 0.5 * torch.randn(B, H, 1, D)).to(dtype)   # a common direction + noise
        if dtype == torch.bfloat16:
            # (in fp16 the reference's eps=1e-12 underflows to 0 and a zero key turns every score into NaN)
            keys[:, :, 7] = 0
        keys[:, :, 20:24] = keys[:, :, 10:14]
        with torch.no_grad():
            sc = kvpress.KeyDiffPress().score(None, None, keys, None, None, {})
        out[f"{tag}_keys"], out[f"{tag}_scores"] = u16(keys), u16(sc)
    np.savez_compressed(OUT / "keydiff.npz", **out)
    print("keydiff", out["cases"].tolist())


def large_inputs(seed: int, B: int, Hkv: int, S: int, D: int, hidden: int, dtype=torch.bfloat16):
    """Seeded inputs of the large case, regenerated (not stored) by the tests: explicit CPU generator, fixed draw
    order. A checksum stored next to the outputs detects an RNG that no longer matches."""
    g = torch.Generator().manual_seed(seed)
    hidden_states = torch.randn(B, S, hidden, generator=g).to(dtype)
    keys = torch.randn(B, Hkv, S, D, generator=g).to(dtype)
    values = torch.randn(B, Hkv, S, D, generator=g).to(dtype)
    return hidden_states, keys, values


def tensor_checksum(*tensors) -> np.ndarray:
    return np.array([int(t.contiguous().view(torch.int16).to(torch.int64).sum()) for t in tensors], dtype=np.int64)


def make_large(name="large32k", **kw):
    out = build_large(**kw)
    np.savez_compressed(OUT / f"{name}.npz", **out)
    print(name, {k: v.shape for k, v in out.items() if k.endswith("scores")})


def build_large(*, B=1, Hq=4, Hkv=1, D=128, hidden=256, S=32768, seed=21):
    """The four in-scope scorers + TOVA at a 32k context, reference run unmodified on CPU. Only the q_proj weight,
    the outputs and an input checksum are stored; K, V and the hidden states are regenerated from the seed."""
    kvpress = import_reference()
    from transformers import LlamaConfig
    from transformers.models.llama.modeling_llama import LlamaAttention, LlamaRotaryEmbedding

    torch.manual_seed(seed)
    cfg = LlamaConfig(hidden_size=hidden, num_attention_heads=Hq, num_key_value_heads=Hkv, head_dim=D,
                      num_hidden_layers=1, intermediate_size=2 * hidden, vocab_size=128,
                      max_position_embeddings=65536, rope_theta=500000.0)
    cfg._attn_implementation = "sdpa"
    attn = LlamaAttention(cfg, 0).to(torch.bfloat16).eval()
    attn.rotary_emb = LlamaRotaryEmbedding(cfg)
    hidden_states, keys, values = large_inputs(seed, B, Hkv, S, D, hidden)
    cos, sin = attn.rotary_emb(hidden_states, torch.arange(S)[None])
    kwargs = {"position_embeddings": (cos, sin)}
    out = {"meta": np.array([B, Hq, Hkv, D, hidden, S, seed], dtype=np.int64),
           "checksum": tensor_checksum(hidden_states, keys, values),
           "q_weight": u16(attn.q_proj.weight.detach()), "rope_theta": np.array([500000.0])}
    ratios = [0.5, 0.7]
    out["ratios"] = np.array(ratios)
    with torch.no_grad():
        presses = {
            "knorm": kvpress.KnormPress(),
            "snap": kvpress.SnapKVPress(window_size=64, kernel_size=5),
            "ea": kvpress.ExpectedAttentionPress(),
            "tova": kvpress.TOVAPress(),
        }
        mu, cov = presses["ea"].get_query_statistics(attn, hidden_states)
        out["ea_mu"], out["ea_cov"] = u16(mu), u16(cov)
        for tag, press in presses.items():
            sc = press.score(attn, hidden_states, keys, values, None, kwargs)
            out[f"{tag}_scores"] = u16(sc)
            for i, r in enumerate(ratios):
                idx = sc.topk(int(S * (1 - r)), dim=-1).indices
                out[f"{tag}_kept_{i}"] = idx.sort(-1).values.numpy().astype(np.int32)
    return out


if __name__ == "__main__":
    if "--large-only" in sys.argv:
        make_large()
        sys.exit(0)
    if "--keydiff-only" in sys.argv:
        make_keydiff()
        sys.exit(0)
    if "--rerotation-only" in sys.argv:
        make_rerotation()
        sys.exit(0)
    for case_name, case_kw in CASES.items():
        make_case(case_name, **case_kw)
    make_decoding_table()
    make_rerotation()
    make_keydiff()
    make_large()
