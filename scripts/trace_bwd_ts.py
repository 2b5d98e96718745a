"""Timeline of one k_l1_bwd_ts launch (products-sized: 65 k rows, 105 columns of Y, D = 256): globaltimer stamps of CTA 0's
MMA issuer (per row group), its epilogue warp (per flush interval) and first converter warp (grapes_tc_debug bits 3 + 4)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grapes_b200._lib import lib, ptr
from grapes_b200.utils import _any_ctx

def main():
    dev = torch.device("cuda", 0)
    L = lib(); holder = _any_ctx(dev); ctx = holder.ctx
    n, K, D = 64943, 104, 256
    ncols, ldy = 105, 108
    Y = torch.randn(n, ldy, device=dev); Y[:, 104] = 1.0
    W = torch.randn(D, K, device=dev) * 0.1
    b1, w2, dz = torch.randn(D, device=dev), torch.randn(D, device=dev), torch.randn(n, device=dev)
    mask = torch.randint(-2**31, 2**31 - 1, (((n + 127) // 128 * 4), D), dtype=torch.int32, device=dev)
    cnt = torch.tensor([n], dtype=torch.int32, device=dev)
    gW, gb, gw = torch.zeros(D, K, device=dev), torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for flags, tag in ((16, "k_l1_bwd_ts"), (8, "k_l1_bwd_tc")):
        L.cdll.grapes_tc_debug(flags)
        def run():
            L.grapes_sampler_l1_bwd_tc(ctx, ptr(Y), None, ldy, ncols, ptr(cnt), n, K, 104, ptr(mask), ptr(W), K, D, ptr(b1), ptr(w2),
                                       ptr(dz), 1.0, ptr(gW), ptr(gb), ptr(gw), st)
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        print(f"{tag}: event time of the launch pair (contraction + finalize): {e0.elapsed_time(e1) * 1e3:.1f} us")
        if flags & 16:
            src = int(L.cdll.grapes_debug_partials(ctx)) + int(L.cdll.grapes_debug_partials_bytes(ctx)) - 1024
            t = torch.empty(96, dtype=torch.int64, device=dev)
            ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(src), 96 * 8, 3)
            v = t.cpu().tolist(); t0 = v[0]
            rel = lambda x: round((x - t0) / 1e3, 2)
            print("  MMA issuer: groups committed", [rel(x) for x in v[2:32] if x])
            print("  epilogue warp 2: flush intervals stored", [rel(x) for x in v[40:48] if x])
            print("  converter warp 6: its groups stored", [rel(x) for x in v[64:80] if x])
    L.cdll.grapes_tc_debug(0)

if __name__ == "__main__":
    main()
