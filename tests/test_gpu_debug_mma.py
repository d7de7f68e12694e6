"""Pins the tcgen05 shared-memory descriptor semantics the tensor-core backward relies on
(csrc/backward_tc.cu), with the development probe nvfi_debug_mma_mn: one 128x128x128 TF32 MMA,
A^T from tensor memory, B from the sample-major 128B-swizzled shared-memory tile.

 * K-major B (rows = N, 128-byte rows of 32 K elements, K blocks 16 KB apart, SBO 1 KB): this is
   how dW reads G^T and how every forward layer reads its weights.  Must equal At @ B^T.
 * MN-major B of the same bytes (what a transpose-free dW would need): kind::tf32 only accepts
   MN-major operands in the SWIZZLE_128B_BASE32B layout, and with the plain 128B swizzle the
   hardware returns zeros.  The test documents that (it is the reason for the L2 round trip)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _tf32(x):
    xi = x.contiguous().view(torch.int32)
    return ((xi + 0x1000) & ~0x1FFF).view(torch.float32)


def _run(At, G, lbo, sbo, kstep, layout, mn):
    from nvfi_b200 import _lib as L
    lib = L.load_debug()
    D = torch.full((128, 128), float("nan"), device="cuda")
    rc = lib.nvfi_debug_mma_mn(At.data_ptr(), G.data_ptr(), D.data_ptr(), lbo, sbo, kstep, layout, mn,
                               torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    return D


def test_k_major_b_operand():
    torch.manual_seed(0)
    At = torch.randn(128, 128, device="cuda")
    G = torch.randn(128, 128, device="cuda")
    D = _run(At, G, 1, 64, 32, 2, 0)
    ref = _tf32(At).double() @ _tf32(G).double().t()
    assert float((D.double() - ref).norm() / ref.norm()) < 1e-6


def test_mn_major_b_with_128b_swizzle_is_not_usable():
    torch.manual_seed(1)
    At = torch.randn(128, 128, device="cuda")
    G = torch.randn(128, 128, device="cuda")
    D = _run(At, G, 1024, 64, 1024, 2, 1)
    ref = _tf32(At).double() @ _tf32(G).double()
    assert float((D.double() - ref).norm() / ref.norm()) > 0.5
