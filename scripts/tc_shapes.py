"""Launch shapes of every tcgen05 conv of one BraTS 2x4x128^3 step (CPU only: tta_conv_tc_query with
TTA_TC_DEBUG=1 prints tiles, split-K, work items and waves)."""
import ctypes, os, sys
os.environ["TTA_TC_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_tta_b200 import _lib
lib = _lib.lib()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
F, H = 0, 2   # TTA_F16 (split planes), TTA_F16_HI (one plane)
L = [  # name, dtype, cin, cout, in size, mode, K, stride
    ("stem S2 4->64", F, 4, 64, S, 0, 3, 2), ("S1 32 @64", F, 32, 32, S // 2, 0, 3, 1),
    ("S2 32->128", F, 32, 128, S // 2, 0, 3, 2), ("S1 64 @32", F, 64, 64, S // 4, 0, 3, 1),
    ("S2 64->256", F, 64, 256, S // 4, 0, 3, 2), ("S1 128 @16", F, 128, 128, S // 8, 0, 3, 1),
    ("S2 128->512", F, 128, 512, S // 8, 0, 3, 2), ("S1 256 @8", F, 256, 256, S // 16, 0, 3, 1),
    ("K1 256->512", F, 256, 512, S // 16, 0, 1, 1), ("S1 256->512 @8", F, 256, 512, S // 16, 0, 3, 1),
    ("S1 512 @8", F, 512, 512, S // 16, 0, 3, 1), ("T2 768->128", F, 768, 128, S // 16, 1, 3, 2),
    ("S1 128 @16", F, 128, 128, S // 8, 0, 3, 1), ("T2 256->64", F, 256, 64, S // 8, 1, 3, 2),
    ("S1 64 @32", F, 64, 64, S // 4, 0, 3, 1), ("T2 128->32", F, 128, 32, S // 4, 1, 3, 2),
    ("S1 32 @64", F, 32, 32, S // 2, 0, 3, 1), ("T2 64->3", F, 64, 3, S // 2, 1, 3, 2),
    ("b S2 3->64", H, 3, 64, S, 0, 3, 2), ("b S1T 32 @64", H, 32, 32, S // 2, 1, 3, 1),
    ("b S2 32->128", H, 32, 128, S // 2, 0, 3, 2), ("b S1T 64 @32", H, 64, 64, S // 4, 1, 3, 1),
    ("b S2 64->256", H, 64, 256, S // 4, 0, 3, 2), ("b S1T 128 @16", H, 128, 128, S // 8, 1, 3, 1),
    ("b S2 128->768", H, 128, 768, S // 8, 0, 3, 2), ("b S1T 512 @8", H, 512, 512, S // 16, 1, 3, 1),
    ("b S1T 512->256 @8", H, 512, 256, S // 16, 1, 3, 1), ("b K1 512->256", H, 512, 256, S // 16, 1, 1, 1),
    ("b S1T 256 @8", H, 256, 256, S // 16, 1, 3, 1), ("b T2 512->128", H, 512, 128, S // 16, 1, 3, 2),
    ("b S1T 128 @16", H, 128, 128, S // 8, 1, 3, 1), ("b T2 256->64", H, 256, 64, S // 8, 1, 3, 2),
    ("b S1T 64 @32", H, 64, 64, S // 4, 1, 3, 1), ("b T2 128->32", H, 128, 32, S // 4, 1, 3, 2),
    ("b S1T 32 @64", H, 32, 32, S // 2, 1, 3, 1),
]
for name, dt, cin, cout, s, mode, K, st in L:
    so = s if st == 1 else (s * 2 if mode == 1 else s // 2)
    ks, grid, nbuf = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    sys.stderr.write(f"{name:20s} ")
    sys.stderr.flush()
    lib.tta_conv_tc_query(dt, N, (cin + 7) // 8, s, s, s, (cout + 7) // 8, so, so, so, mode, K, st, 0, 0,
                          ctypes.byref(ks), ctypes.byref(grid), ctypes.byref(nbuf))
