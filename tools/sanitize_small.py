"""Small end-to-end calls of every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvpress_b200 import native
from oracle import press_oracle as O
torch.manual_seed(0)
dev = "cuda:0"
B, H, Hq, S, D, w = 1, 2, 8, 1500, 128, 64
k = torch.randn(B, H, S, D, dtype=torch.bfloat16, device=dev); v = torch.randn_like(k)
q = torch.randn(B, Hq, w, D, dtype=torch.bfloat16, device=dev)
mu = (0.5 * torch.randn(B, Hq, D, device=dev)).to(torch.bfloat16)
a = torch.randn(B, Hq, D, D, device=dev) / D ** 0.5
cov = (a @ a.transpose(-1, -2)).to(torch.bfloat16)
n_kept = 700
outs = []
outs.append(native.knorm_compress(k, v, n_kept, return_indices=True, return_scores=True))
big_k = torch.randn(1, 8, 20000, 128, dtype=torch.bfloat16, device=dev)  # > 32 MiB -> two-kernel Knorm path
outs.append(native.knorm_compress(big_k, big_k, 9000, return_indices=True))
outs.append(native.streaming_compress(k, v, n_kept, 4, return_indices=True))
outs.append(native.snapkv_compress(k, v, q, w, 5, n_kept, return_indices=True, return_scores=True))
outs.append(native.expected_attention_compress(k, v, mu, cov, 0.0, 4, True, n_kept, return_indices=True, return_scores=True))
outs.append(native.expected_attention_compress(k, v, mu, None, 0.0, 4, True, n_kept, return_indices=True))
# group sizes 2 (CTA-pair kernel, one half per tile) and 3 (one-CTA kernel with a padding head)
for hq in (4, 6):
    outs.append(native.expected_attention_compress(k, v, mu[:, :hq].contiguous(), cov[:, :hq].contiguous(), 0.0, 4, True, n_kept,
                                                   return_indices=True))
sc = torch.randn(B, H, S, device=dev).to(torch.bfloat16)
outs.append(native.scores_compress(sc, k, v, n_kept, return_indices=True))
torch.cuda.synchronize()
idx = outs[0][2].cpu().long()
assert torch.equal(outs[0][0].cpu(), O.gather_rows(k.cpu(), idx))
print("sanitize_small: all kernels ran")
