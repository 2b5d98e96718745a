"""tcgen05 (3xTF32) sampler-layer kernel against a float64 torch reference of the same op:
z = relu(Y W1^T + b1) . w2, tolerance 1e-5 relative (the fp32 bar of BASELINE.json)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _split(Y, presplit, dev):
    """(Y, None): split inside the kernels; (Y_hi, Y_lo): the pair the aggregation writes."""
    if not presplit:
        return Y, None
    from grapes_b200._lib import lib, ptr
    from grapes_b200.utils import _any_ctx
    Yh, Yl = torch.empty_like(Y), torch.empty_like(Y)
    lib().grapes_split_tf32(_any_ctx(dev).ctx, ptr(Y), Y.shape[1], Y.shape[0], Y.shape[1], ptr(Yh), ptr(Yl), Y.shape[1],
                            torch.cuda.current_stream().cuda_stream)
    return Yh, Yl


def _run_tc(Y, W1, b1, w2, n, dev, with_mask=False, presplit=True):
    from grapes_b200._lib import lib, ptr
    from grapes_b200.utils import _any_ctx
    L, ctx = lib(), _any_ctx(dev).ctx
    st = torch.cuda.current_stream().cuda_stream
    cap_n, ldy = Y.shape
    D, K = W1.shape
    ldw = (K + 3) // 4 * 4
    Wh = torch.empty((D, ldw), device=dev); Wl = torch.empty((D, ldw), device=dev)
    L.grapes_split_tf32(ctx, ptr(W1), K, D, K, ptr(Wh), ptr(Wl), ldw, st)
    zpart = torch.zeros((2 * (D // 128), cap_n), device=dev)
    cnt = torch.tensor([n], dtype=torch.int32, device=dev)
    maskT = torch.zeros(((cap_n + 127) // 128 * 4, D), dtype=torch.int32, device=dev) if with_mask else None
    Ya, Yb = _split(Y, presplit, dev)
    L.grapes_sampler_l1_fwd_tc(ctx, ptr(Ya), ptr(Yb), ldy, ptr(cnt), cap_n, K, ptr(Wh), ptr(Wl), ldw, D, ptr(b1),
                               ptr(w2), ptr(zpart), ptr(maskT), st)
    torch.cuda.synchronize()
    return zpart.sum(0)[:n], maskT


@pytest.mark.parametrize("presplit", [True, False])
@pytest.mark.parametrize("n,cap_n,K,D", [(1000, 1500, 104, 256), (64943, 66000, 104, 256), (130, 256, 15, 128),
                                         (5000, 5000, 605, 256), (1, 128, 100, 256), (777, 900, 1436, 384),
                                         (3000, 3000, 131, 256)])
def test_l1_fwd_tc_matches_fp64(cuda_device, n, cap_n, K, D, presplit):
    g = torch.Generator().manual_seed(n + K)
    ldy = (K + 3) // 4 * 4
    Y = torch.zeros(cap_n, ldy)
    Y[:, :K] = torch.randn(cap_n, K, generator=g)
    W1 = (torch.rand(D, K, generator=g) * 2 - 1) * (6.0 / (D + K)) ** 0.5
    b1 = torch.randn(D, generator=g) * 0.1
    w2 = torch.randn(D, generator=g) * 0.1
    ref = (torch.relu(Y[:n, :K].double() @ W1.double().t() + b1.double()) * w2.double()).sum(1)
    got, _ = _run_tc(Y.to(cuda_device), W1.to(cuda_device), b1.to(cuda_device), w2.to(cuda_device), n, cuda_device,
                     presplit=presplit)
    err = (got.double().cpu() - ref).abs().max() / ref.abs().max()
    assert err < 1e-5, f"relative error {err:.3e}"


def test_l1_fwd_tc_relu_mask_bits(cuda_device):
    n, cap_n, K, D = 3000, 3072, 104, 256
    g = torch.Generator().manual_seed(1)
    Y = torch.randn(cap_n, K, generator=g)
    W1 = torch.randn(D, K, generator=g) * 0.1
    b1 = torch.randn(D, generator=g)
    w2 = torch.randn(D, generator=g)
    pre = Y[:n].double() @ W1.double().t() + b1.double()
    _, maskT = _run_tc(Y.to(cuda_device), W1.to(cuda_device), b1.to(cuda_device), w2.to(cuda_device), n, cuda_device,
                       with_mask=True)
    m = maskT.cpu()
    rows = torch.arange(n)
    bits = (m[rows // 32] >> (rows % 32).unsqueeze(1)) & 1          # [n, D]
    want = (pre > 0)
    clear = pre.abs() > 1e-4                                        # away from the relu kink the mask is exact
    assert torch.equal(bits.bool()[clear], want[clear])
    assert (bits.bool() != want).float().mean() < 1e-4


@pytest.mark.parametrize("presplit", [True, False])
@pytest.mark.parametrize("n,cap_n,K,D,extra", [(3000, 3072, 104, 256, 0), (64943, 66000, 104, 256, 0), (100, 128, 15, 128, 0),
                                               (2000, 2048, 100, 256, 4), (5000, 5100, 131, 256, 0),
                                               (4000, 4096, 605, 256, 0), (1500, 1536, 1433, 256, 3)])
def test_l1_bwd_tc_matches_fp64_autograd(cuda_device, n, cap_n, K, D, extra, presplit):
    """d(sum_r dz[r] z[r]) / d(W1, b1, w2) from the relu-mask bits: S = mask^T (dz * [Y | 1]).
    `extra` indicator-like columns sit between K and the ones column (the gcn_z case, K = F < F')."""
    from grapes_b200._lib import lib, ptr
    from grapes_b200.utils import _any_ctx
    dev = cuda_device
    L, ctx = lib(), _any_ctx(dev).ctx
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(n + K + extra)
    ones_col = K + extra
    ncols = ones_col + 1
    ldy = (ncols + 3) // 4 * 4
    Y = torch.zeros(cap_n, ldy)
    Y[:, :ones_col] = torch.randn(cap_n, ones_col, generator=g)
    Y[:, ones_col] = 1.0
    W1 = (torch.rand(D, K, generator=g) * 2 - 1) * (6.0 / (D + K)) ** 0.5
    b1 = torch.randn(D, generator=g) * 0.1
    w2 = torch.randn(D, generator=g) * 0.1
    dz = torch.randn(cap_n, generator=g)
    pre64 = Y[:n, :K].double() @ W1.double().t() + b1.double()
    Yd, W1d, b1d, w2d, dzd = (t.to(dev) for t in (Y, W1, b1, w2, dz))
    ldw = (K + 3) // 4 * 4
    Wh, Wl = torch.empty((D, ldw), device=dev), torch.empty((D, ldw), device=dev)
    L.grapes_split_tf32(ctx, ptr(W1d), K, D, K, ptr(Wh), ptr(Wl), ldw, st)
    zpart = torch.zeros((2 * (D // 128), cap_n), device=dev)
    cnt = torch.tensor([n], dtype=torch.int32, device=dev)
    maskT = torch.zeros(((cap_n + 127) // 128 * 4, D), dtype=torch.int32, device=dev)
    Ya, Yb = _split(Yd, presplit, dev)
    L.grapes_sampler_l1_fwd_tc(ctx, ptr(Ya), ptr(Yb), ldy, ptr(cnt), cap_n, K, ptr(Wh), ptr(Wl), ldw, D, ptr(b1d),
                               ptr(w2d), ptr(zpart), ptr(maskT), st)
    gW1, gb1, gw2 = torch.zeros(D, K, device=dev), torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    L.grapes_sampler_l1_bwd_tc(ctx, ptr(Ya), ptr(Yb), ldy, ncols, ptr(cnt), cap_n, K, ones_col, ptr(maskT), ptr(W1d), K,
                               D, ptr(b1d), ptr(w2d), ptr(dzd), 1.0, ptr(gW1), ptr(gb1), ptr(gw2), st)
    torch.cuda.synchronize()
    # the relu mask the forward emitted: identical to float64's away from the kink; AT the kink (|pre| within fp32
    # noise of 0) either branch is a valid fp32 answer, so the float64 reference is evaluated with the kernel's choice
    rows = torch.arange(n)
    bits = ((maskT.cpu()[rows // 32] >> (rows % 32).unsqueeze(1)) & 1).bool()
    want = pre64 > 0
    assert torch.equal(bits[pre64.abs() > 1e-5], want[pre64.abs() > 1e-5])
    assert (bits != want).sum() <= 4
    m = bits.double()
    g = dz[:n].double().unsqueeze(1) * m * w2.double()                       # d(sum dz z)/d pre
    refs = ((gW1, g.t() @ Y[:n, :K].double(), "W1"), (gb1, g.sum(0), "b1"),
            (gw2, (dz[:n].double().unsqueeze(1) * m * pre64).sum(0), "w2"))
    for got, ref, name in refs:
        err = (got.double().cpu() - ref).abs().max() / ref.abs().max()
        assert err < 1e-5, f"{name}: relative error {err:.3e}"


@pytest.mark.parametrize("n,cap_n,K,D", [(1000, 1500, 104, 256), (64943, 66000, 104, 256), (130, 256, 15, 128),
                                         (5000, 5000, 605, 256), (1, 128, 100, 256), (777, 900, 1436, 384),
                                         (3000, 3000, 131, 256), (40000, 40001, 100, 256)])
def test_l1_fwd_ts_tensor_memory_operand(cuda_device, n, cap_n, K, D):
    """k_l1_fwd_ts (default forward: raw Y tile -> TMA ring -> registers -> tensor memory, the MMA takes its A operand
    from TMEM) against float64 at 1e-5, and against k_l1_fwd_tc (both operands in shared memory, grapes_tc_debug bit 2):
    same MMA order and epilogue; the lo part of the split is left unrounded here (the tensor core drops its low bits), so
    z agrees to fp32 rounding and the relu mask away from the kink.  Stale rows behind n and columns behind K
    (indicator-like values) must not leak."""
    from grapes_b200._lib import lib
    g = torch.Generator().manual_seed(n + K)
    ldy = (K + 1 + 3) // 4 * 4
    Y = torch.randn(cap_n, ldy, generator=g)                      # pad columns hold finite junk, like the ones column
    W1 = (torch.rand(D, K, generator=g) * 2 - 1) * (6.0 / (D + K)) ** 0.5
    b1 = torch.randn(D, generator=g) * 0.1
    w2 = torch.randn(D, generator=g) * 0.1
    ref = (torch.relu(Y[:n, :K].double() @ W1.double().t() + b1.double()) * w2.double()).sum(1)
    args = (Y.to(cuda_device), W1.to(cuda_device), b1.to(cuda_device), w2.to(cuda_device), n, cuda_device)
    z_ts, m_ts = _run_tc(*args, with_mask=True, presplit=False)
    try:
        lib().cdll.grapes_tc_debug(4)
        z_tc, m_tc = _run_tc(*args, with_mask=True, presplit=False)
    finally:
        lib().cdll.grapes_tc_debug(0)
    err = (z_ts.double().cpu() - ref).abs().max() / ref.abs().max()
    assert err < 1e-5, f"relative error {err:.3e}"
    assert ((z_ts - z_tc).abs().max() / z_tc.abs().max()).item() < 2e-6
    rows = torch.arange(n)
    bits = [((m.cpu()[rows // 32] >> (rows % 32).unsqueeze(1)) & 1).bool() for m in (m_ts, m_tc)]
    pre = Y[:n, :K].double() @ W1.double().t() + b1.double()
    clear = pre.abs() > 1e-4 * pre.abs().max()
    assert torch.equal(bits[0][clear], bits[1][clear]) and torch.equal(bits[0][clear], (pre > 0)[clear])
    assert (bits[0] != bits[1]).sum() <= 4


@pytest.mark.parametrize("n,cap_n,K,D,extra", [(3000, 3072, 104, 256, 0), (64943, 66000, 104, 256, 0), (100, 128, 15, 128, 0),
                                               (2000, 2048, 100, 256, 4), (5000, 5100, 131, 256, 0),
                                               (4000, 4096, 605, 256, 0), (1500, 1536, 1433, 256, 3)])
def test_l1_bwd_ts_tensor_memory_operand(cuda_device, n, cap_n, K, D, extra):
    """k_l1_bwd_ts (S^T = (dz * Y)^T mask with the scaled, split features in tensor memory, one hidden half per CTA) against
    float64 autograd at 1e-5 (evaluated with the kernel's own relu mask, as in test_l1_bwd_tc_matches_fp64_autograd) and
    against k_l1_bwd_tc to fp32 rounding (the row groups are dealt to the CTAs differently, so not bit for bit)."""
    from grapes_b200._lib import lib, ptr
    from grapes_b200.utils import _any_ctx
    dev = cuda_device
    L, ctx = lib(), _any_ctx(dev).ctx
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(n + K + extra)
    ones_col = K + extra
    ncols = ones_col + 1
    ldy = (ncols + 3) // 4 * 4
    Y = torch.zeros(cap_n, ldy)
    Y[:, :ones_col] = torch.randn(cap_n, ones_col, generator=g)
    Y[:, ones_col] = 1.0
    W1 = (torch.rand(D, K, generator=g) * 2 - 1) * (6.0 / (D + K)) ** 0.5
    b1 = torch.randn(D, generator=g) * 0.1
    w2 = torch.randn(D, generator=g) * 0.1
    dz = torch.randn(cap_n, generator=g)
    pre64 = Y[:n, :K].double() @ W1.double().t() + b1.double()
    Yd, W1d, b1d, w2d, dzd = (t.to(dev) for t in (Y, W1, b1, w2, dz))
    ldw = (K + 3) // 4 * 4
    Wh, Wl = torch.empty((D, ldw), device=dev), torch.empty((D, ldw), device=dev)
    L.grapes_split_tf32(ctx, ptr(W1d), K, D, K, ptr(Wh), ptr(Wl), ldw, st)
    zpart = torch.zeros((2 * (D // 128), cap_n), device=dev)
    cnt = torch.tensor([n], dtype=torch.int32, device=dev)
    maskT = torch.zeros(((cap_n + 127) // 128 * 4, D), dtype=torch.int32, device=dev)
    L.grapes_sampler_l1_fwd_tc(ctx, ptr(Yd), None, ldy, ptr(cnt), cap_n, K, ptr(Wh), ptr(Wl), ldw, D, ptr(b1d),
                               ptr(w2d), ptr(zpart), ptr(maskT), st)
    outs = []
    try:
        for flag in (0, 8):                      # 0: k_l1_bwd_ts (default), 8: k_l1_bwd_tc
            L.cdll.grapes_tc_debug(flag)
            gW1, gb1, gw2 = torch.zeros(D, K, device=dev), torch.zeros(D, device=dev), torch.zeros(D, device=dev)
            L.grapes_sampler_l1_bwd_tc(ctx, ptr(Yd), None, ldy, ncols, ptr(cnt), cap_n, K, ones_col, ptr(maskT), ptr(W1d), K,
                                       D, ptr(b1d), ptr(w2d), ptr(dzd), 1.0, ptr(gW1), ptr(gb1), ptr(gw2), st)
            torch.cuda.synchronize()
            outs.append((gW1, gb1, gw2))
    finally:
        L.cdll.grapes_tc_debug(0)
    rows = torch.arange(n)
    m = ((maskT.cpu()[rows // 32] >> (rows % 32).unsqueeze(1)) & 1).double()
    gg = dz[:n].double().unsqueeze(1) * m * w2.double()
    refs = (gg.t() @ Y[:n, :K].double(), gg.sum(0), (dz[:n].double().unsqueeze(1) * m * pre64).sum(0))
    for got, old, ref, name in zip(outs[0], outs[1], refs, ("W1", "b1", "w2")):
        err = (got.double().cpu() - ref).abs().max() / ref.abs().max()
        assert err < 1e-5, f"{name}: relative error {err:.3e} against float64"
        dif = (got.double() - old.double()).abs().max().item() / ref.abs().max().item()
        assert dif < 2e-6, f"{name}: {dif:.3e} away from k_l1_bwd_tc"
