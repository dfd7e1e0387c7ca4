import sys; sys.path.insert(0, '/root/repo')
import torch
from kvpress_b200 import native
from oracle import press_oracle as O
for shape, nk in [((2,2,384,64), 192), ((1,8,2560,128), 2048), ((1,2,300,128), 100), ((1,1,256,128), 128), ((1,1,257,128), 128)]:
    torch.manual_seed(0)
    k = torch.randn(shape, dtype=torch.bfloat16, device="cuda"); v = torch.randn_like(k)
    ko, vo, idx, sc = native.knorm_compress(k, v, nk, return_indices=True, return_scores=True)
    torch.cuda.synchronize()
    want = O.select_lowest_index_ties(sc.cpu(), nk)
    print(shape, nk, "idx ok:", torch.equal(idx.cpu().long(), want), "scores ok:", torch.equal(sc.cpu(), O.knorm_scores(k.cpu())), idx[0,0,:6].tolist(), want[0,0,:6].tolist())
