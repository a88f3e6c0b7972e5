"""GPU test of the tcgen05 / TMA / TMEM primitives the tile kernel is assembled from."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_umma_selftest_matches_numpy():
    from pytorch_simclr_b200 import _lib
    lib = _lib.load_debug()      # the self-test kernel lives in the tracing build only
    gen = torch.Generator().manual_seed(0)
    a = torch.randn(128, 128, generator=gen).to(torch.bfloat16).cuda()
    b = torch.randn(128, 128, generator=gen).to(torch.bfloat16).cuda()
    out = torch.full((3, 128, 128), float("nan"), device="cuda")
    _lib.check(lib.simclr_selftest_umma(a.data_ptr(), b.data_ptr(), out.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream), "selftest")
    torch.cuda.synchronize()
    af, bf = a.float().cpu().double().numpy(), b.float().cpu().double().numpy()
    o = out.cpu().double().numpy()
    # fp32 accumulation of exact bf16 products: error is accumulation round-off only
    assert np.abs(o[0] - af @ bf.T).max() < 1e-4      # SS, both K-major (TMA 128B swizzle)
    assert np.abs(o[1] - af @ bf).max() < 1e-4        # SS, B read as MN-major
    assert np.abs(o[2] - af @ bf).max() < 1e-4        # TS, A written to TMEM with tcgen05.st
    assert np.array_equal(o[1], o[2])                 # same products, same order
