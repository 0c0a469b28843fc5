import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, makb200
from oracle import mak_oracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
A = makb200.to_device(O.rand_hermitian(n, "f64", 1))
D, V = makb200.eigh_full(A)
torch.cuda.synchronize()
print("ok", float(np.abs(D.cpu().numpy() - np.linalg.eigvalsh(O.rand_hermitian(n, "f64", 1))).max()))
