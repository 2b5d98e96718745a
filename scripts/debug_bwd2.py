import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from grapes_b200._lib import lib, ptr
from grapes_b200.utils import _any_ctx
dev = torch.device("cuda:0")
L, ctx = lib(), _any_ctx(dev).ctx
st = torch.cuda.current_stream().cuda_stream
n, cap_n, K, D = int(sys.argv[1]), int(sys.argv[2]), 104, 256
g = torch.Generator().manual_seed(0)
ones_col = K; ncols = K + 1; ldy = 108
Y = torch.zeros(cap_n, ldy); Y[:, :K] = torch.randn(cap_n, K, generator=g); Y[:, ones_col] = 1.0
W1 = torch.randn(D, K, generator=g) * 0.2; b1 = torch.randn(D, generator=g) * 0.1; w2 = torch.randn(D, generator=g)
dz = torch.randn(cap_n, generator=g)
pre = Y[:n, :K].double() @ W1.double().t() + b1.double()
mask = (pre > 0)
S_ref = (mask.double() * dz[:n].double().unsqueeze(1)).t() @ Y[:n, :ncols].double()
Yd, W1d, b1d, w2d, dzd = (t.to(dev) for t in (Y, W1, b1, w2, dz))
ldw = 104
Yh, Yl = torch.empty_like(Yd), torch.empty_like(Yd)
Wh, Wl = torch.empty((D, ldw), device=dev), torch.empty((D, ldw), device=dev)
L.grapes_split_tf32(ctx, ptr(Yd), ldy, cap_n, ldy, ptr(Yh), ptr(Yl), ldy, st)
L.grapes_split_tf32(ctx, ptr(W1d), K, D, K, ptr(Wh), ptr(Wl), ldw, st)
zpart = torch.zeros((D // 128, cap_n), device=dev)
cnt = torch.tensor([n], dtype=torch.int32, device=dev)
ng = (cap_n + 127) // 128 * 4
maskT = torch.zeros((ng, D), dtype=torch.int32, device=dev)
L.grapes_sampler_l1_fwd_tc(ctx, ptr(Yh), ptr(Yl), ldy, ptr(cnt), cap_n, K, ptr(Wh), ptr(Wl), ldw, D, ptr(b1d), ptr(w2d), ptr(zpart), ptr(maskT), st)
torch.cuda.synchronize()
m = maskT.cpu()
rows = torch.arange(n)
bits = ((m[rows // 32] >> (rows % 32).unsqueeze(1)) & 1).bool()
bad = (bits != mask)
print("mask mismatches:", int(bad.sum()), "of", bad.numel(), " near-kink mismatches:", int((bad & (pre.abs() < 1e-4)).sum()))
if bad.any():
    idx = bad.nonzero()[:10]; print("first bad (row, d):", idx.tolist(), "pre:", [float(pre[i, j]) for i, j in idx.tolist()])
    print("bad rows histogram by (row//128):", torch.bincount(bad.nonzero()[:, 0] // 128)[:20].tolist())
# host-built mask words -> bwd
pad = torch.zeros(ng * 32, D, dtype=torch.bool); pad[:n] = mask
w = (pad.view(ng, 32, D).long() << torch.arange(32).view(1, 32, 1)).sum(1)
w = torch.where(w >= 2**31, w - 2**32, w).to(torch.int32).to(dev)
for name, mk in (("device mask", maskT), ("host mask", w)):
    gW1, gb1, gw2 = torch.zeros(D, K, device=dev), torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    L.grapes_sampler_l1_bwd_tc(ctx, ptr(Yh), ptr(Yl), ldy, ncols, ptr(cnt), cap_n, K, ones_col, ptr(mk), ptr(W1d), K, D, ptr(b1d), ptr(w2d), ptr(dzd), 1.0, ptr(gW1), ptr(gb1), ptr(gw2), st)
    torch.cuda.synchronize()
    S_got = gW1.double().cpu() / w2.double().unsqueeze(1)
    err = (S_got - S_ref[:, :K]).abs()
    print(name, "S rel err", (err.max() / S_ref.abs().max()).item(), "argmax", divmod(int(err.argmax()), K))
    print("   per-d-half max err:", err[:128].max().item(), err[128:].max().item(), " cols>=96 err", err[:, 96:].max().item(), "cols<32", err[:, :32].max().item())
