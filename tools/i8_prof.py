"""One int8 Gram (and one fp64 Gram) for ncu launch lists: python tools/i8_prof.py n k [path]"""
import sys
import torch
sys.path.insert(0, ".")
from fitsnap_b200.engine import Engine
n, k = int(sys.argv[1]), int(sys.argv[2])
path = sys.argv[3] if len(sys.argv) > 3 else "int8"
eng = Engine(0)
g = torch.Generator(device=eng.device); g.manual_seed(1)
A = torch.randn((n, k), dtype=torch.float64, device=eng.device, generator=g)
b = torch.randn(n, dtype=torch.float64, device=eng.device, generator=g)
w = torch.ones(n, dtype=torch.float64, device=eng.device)
eng.set_gram_path(path)
for _ in range(2):
    G = eng.gram(A, b, w)
torch.cuda.synchronize()
