"""Timeline of one k_l1_fwd_ts launch (products-sized: 65 k rows, K = 104, D = 256): globaltimer stamps of CTA 0's MMA issuer,
epilogue warp and first converter warp (grapes_tc_debug bit 4 parks them in the ctx's partial buffer)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grapes_b200._lib import lib, ptr
from grapes_b200.utils import _any_ctx

def main():
    dev = torch.device("cuda", 0)
    L = lib(); holder = _any_ctx(dev); ctx = holder.ctx
    n, K, D = 64943, 104, 256
    ldy = 108
    Y = torch.randn(n, ldy, device=dev)
    W = torch.randn(D, K, device=dev) * 0.1
    Wh, Wl = torch.empty(D, 104, device=dev), torch.empty(D, 104, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    L.grapes_split_tf32(ctx, ptr(W), K, D, K, ptr(Wh), ptr(Wl), 104, st)
    b1, w2 = torch.randn(D, device=dev), torch.randn(D, device=dev)
    zpart = torch.zeros(4, n, device=dev)
    mask = torch.zeros(((n + 127) // 128 * 4, D), dtype=torch.int32, device=dev)
    cnt = torch.tensor([n], dtype=torch.int32, device=dev)
    L.cdll.grapes_tc_debug(16)
    use_mask = "--no-mask" not in sys.argv
    def run():
        L.grapes_sampler_l1_fwd_tc(ctx, ptr(Y), None, ldy, ptr(cnt), n, K, ptr(Wh), ptr(Wl), 104, D, ptr(b1), ptr(w2), ptr(zpart),
                                   ptr(mask) if use_mask else None, st)
    print("relu mask output:", use_mask)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    # the stamps live at the start of the ctx's partial buffer: read through a debug copy kernel-free path (cudaMemcpy)
    buf = (ctypes.c_uint64 * 96)()
    cudart = ctypes.CDLL("libcudart.so")
    part = ctypes.c_void_p()
    # partial buffer address: exported for debugging
    src = int(L.cdll.grapes_debug_partials(ctx))
    t = torch.empty(96, dtype=torch.int64, device=dev)
    cudart.cudaMemcpy(ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(src), 96 * 8, 3)
    v = t.cpu().tolist()
    t0 = v[0]
    rel = lambda x: round((x - t0) / 1e3, 2)
    print(f"event time of the launch: {e0.elapsed_time(e1) * 1e3:.1f} us")
    print("MMA issuer: start 0, W resident", rel(v[1]), "tiles committed", [rel(x) for x in v[2:12] if x])
    print("epilogue warp 2: tiles drained", [rel(x) for x in v[32:42] if x])
    print("converter warp 6: k-blocks stored", [rel(x) for x in v[64:80] if x])
    L.cdll.grapes_tc_debug(0)

if __name__ == "__main__":
    main()
