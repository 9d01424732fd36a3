"""Per-layer conv timing on the GPU (CUDA events, L2 flushed between launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_tta_b200 import _lib
from multimodal_tta_b200._lib import TTA_F16, TTA_BF16, check
from multimodal_tta_b200.layout import pack_weights_tc, wg_forward

lib = _lib.lib()
dev = torch.device("cuda")
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
only = sys.argv[2] if len(sys.argv) > 2 else ""
N = 2
# name, cin, cout, K, stride, mode(0 conv/1 convT-like), input dims
LAYERS = [
    ("stem 4->32 s2 (S2)", 4, 32, 3, 2, 0, 128),
    ("32->32 s1 @64 (S1)", 32, 32, 3, 1, 0, 64),
    ("32->64 s2 (S2)", 32, 64, 3, 2, 0, 64),
    ("64->64 s1 @32 (S1)", 64, 64, 3, 1, 0, 32),
    ("64->128 s2 (S2)", 64, 128, 3, 2, 0, 32),
    ("128->128 s1 @16", 128, 128, 3, 1, 0, 16),
    ("128->256 s2", 128, 256, 3, 2, 0, 16),
    ("256->256 s1 @8", 256, 256, 3, 1, 0, 8),
    ("256->512 s1 @8", 256, 512, 3, 1, 0, 8),
    ("512->512 s1 @8", 512, 512, 3, 1, 0, 8),
    ("256->512 1x1 @8", 256, 512, 1, 1, 0, 8),
    ("768->128 T2 8->16", 768, 128, 3, 2, 1, 8),
    ("256->64 T2 16->32", 256, 64, 3, 2, 1, 16),
    ("128->32 T2 32->64", 128, 32, 3, 2, 1, 32),
    ("64->3 T2 64->128", 64, 3, 3, 2, 1, 64),
    ("3->3 s1 @128", 3, 3, 3, 1, 0, 128),
    ("3->64 s2 128->64 (dgrad head)", 3, 64, 3, 2, 0, 128),
    ("32->128 s2 64->32 (dgrad T2)", 32, 128, 3, 2, 0, 64),
    ("32->4.. skip", 0, 0, 0, 0, 0, 0),
]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
tot = 0.0
for name, cin, cout, K, s, mode, S in LAYERS:
    if cin == 0 or (only and only not in name):
        continue
    c8i, c8o = (cin + 7) // 8, (cout + 7) // 8
    So = S // s if mode == 0 else S * s
    hi = torch.randint(-3000, 3000, (N, c8i, S, S, S, 8), dtype=torch.int16, device=dev)
    lo = torch.zeros_like(hi)
    w = torch.randn(27 if K == 3 else 1, cin, cout, device=dev) * 0.05
    # flags bits 8..10 (real Cout) select the dense-GEMM + col2im kernel where the layer qualifies
    t2s = bool(flags & 0x700) and bool(lib.tta_conv_tc_t2s(mode, K, s, cin, cout, 1))
    lflags = (flags & ~0x700) | ((cout << 8) if t2s else 0)
    wp = pack_weights_tc(w, mode, K, s, TTA_F16, t2s=t2s)
    out = torch.zeros((N, c8o, So, So, So, 8), device=dev)
    args = (hi.data_ptr(), lo.data_ptr(), c8i * S ** 3 * 8, TTA_F16, N, c8i, S, S, S, wp.data_ptr(), 0, out.data_ptr(),
            c8o * So ** 3 * 8, c8o, So, So, So, mode, K, s, 0, lflags)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        check(lib.tta_conv_tc(*args, 0, 0, st))
    ts = []
    for _ in range(5):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); check(lib.tta_conv_tc(*args, 0, 0, st)); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    us = sorted(ts)[2]
    Vout = So ** 3 if mode == 0 else S ** 3 * 8
    flops = 2.0 * N * (So ** 3 if mode == 0 else S ** 3) * (K ** 3) * cin * cout
    gb = N * (c8i * S ** 3 * 32 + c8o * So ** 3 * 32) / 1e9
    tot += us
    print(f"{name:32s} {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s  {gb / us * 1e6:7.0f} GB/s(alg)")
print(f"total {tot:.1f} us")
