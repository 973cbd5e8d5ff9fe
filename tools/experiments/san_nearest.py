import sys, torch
sys.path.insert(0, '.')
from quip_for_all_b200 import codebook_id, _native
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for name in ("E8P12", "E8P12RVQ4B", "E8P12RVQ3B"):
    cb = codebook_id[name](inference=False).to(dev)
    for m in (3, 33, 700):
        x = (torch.randn(m, 8, generator=g) * 1.1).to(dev)
        v, i = cb.quantize(x)
        torch.cuda.synchronize()
        print("nearest(struct)", name, m, int(i.max()))
