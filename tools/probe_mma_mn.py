"""Probe the MN-major shared-memory descriptor semantics of tcgen05.mma kind::tf32."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nvfi_b200 import _lib as L
lib = L.load()
torch.manual_seed(0)
def tf32(x):
    xi = x.contiguous().view(torch.int32)
    return ((xi + 0x1000) & ~0x1FFF).view(torch.float32)
At = torch.randn(128, 128, device="cuda"); G = torch.randn(128, 128, device="cuda")
ref = (tf32(At).double() @ tf32(G).double())
refT = (tf32(At).double() @ tf32(G).double().t())
s = torch.cuda.current_stream().cuda_stream
for name, lbo, sbo, kstep, lt, mn in [("lbo=16K sbo=1K kstep=1K SW128 MN", 1024, 64, 1024, 2, 1),
                                      ("lbo=1K sbo=16K kstep=1K SW128 MN", 64, 1024, 1024, 2, 1),
                                      ("lbo=16K sbo=1K kstep=128 SW128 MN", 1024, 64, 128, 2, 1),
                                      ("lbo=1 sbo=1K kstep=32 SW128 K-major (G^T)", 1, 64, 32, 2, 0)]:
    D = torch.full((128, 128), float("nan"), device="cuda")
    rc = lib.nvfi_debug_mma_mn(At.data_ptr(), G.data_ptr(), D.data_ptr(), lbo, sbo, kstep, lt, mn, s)
    torch.cuda.synchronize()
    e = float((D.double() - ref).norm() / ref.norm())
    eT = float((D.double() - refT).norm() / refT.norm())
    print(f"{name:48s} rc={rc} |D|={float(D.norm()):.3e} err vs A.G {e:.3e}  vs A.G^T {eT:.3e}")
