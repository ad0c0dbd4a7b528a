#!/usr/bin/env python
"""Pin the shared-memory layout tcgen05.mma reads for an MN-major tf32 operand (vihmc_debug_umma).

A is a K-major 'selector' (A[m, k] = 1 iff m == k, m < 8) in the layout the production kernel already uses; B's 8 KB image
holds B_img[i] = i.  Then D[k, n] = the index of the shared-memory word the tensor core read as B[k, n]."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vi-hmc_b200")]
import numpy as np, torch
from vihmc import _lib

dev = torch.device("cuda:0")
lib = _lib.load()

def kmajor_image(M):            # M[128, 8] -> K-major no-swizzle image: (r/8)*SBO + (k/4)*LBO + (r%8)*16 + (k%4)*4, LBO 128 SBO 512
    img = np.zeros(2048, np.float32)
    for r in range(128):
        for k in range(8):
            img[((r // 8) * 512 + (k // 4) * 128 + (r % 8) * 16 + (k % 4) * 4) // 4] = M[r, k]
    return img

def run(a_img, b_img, a_lbo, a_sbo, b_lbo, b_sbo, extra, a_type=0, b_type=0):
    a = torch.from_numpy(a_img).to(dev); b = torch.from_numpy(b_img).to(dev)
    out = torch.full((128, 128), -7.0, device=dev)
    _lib.check(lib.vihmc_debug_umma(a.data_ptr(), b.data_ptr(), a_lbo, a_sbo, b_lbo, b_sbo, a_type, b_type, extra, out.data_ptr(), None))
    torch.cuda.synchronize()
    return out.cpu().numpy()

sel = np.zeros((128, 8), np.float32)
for k in range(8): sel[k, k] = 1.0
ramp = np.arange(2048, dtype=np.float32)
d = run(kmajor_image(sel), ramp, 128, 512, 128, 512, 0)
print("K-major B sanity:", d[:2, :9].astype(int).tolist())
np.set_printoptions(linewidth=250)
# one configuration per process (a bad stride faults the context): python tools/probe_umma_layout.py B|A type lbo sbo
which, bt, lbo, sbo = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
if which == "B":
    d = run(kmajor_image(sel), ramp, 128, 512, lbo, sbo, 1 << 16, 0, bt)
    print(f"MN-major B type {bt} LBO={lbo} SBO={sbo}: word read for (k rows 0..7, n cols 0..39)"); print(d[:8, :40].astype(int))
    print(" n = 64..71, 120..127:"); print(d[:8, 64:72].astype(int)); print(d[:8, 120:128].astype(int))
else:
    d = run(ramp, kmajor_image(sel), lbo, sbo, 128, 512, 1 << 15, bt, 0)
    print(f"MN-major A type {bt} LBO={lbo} SBO={sbo}: word read for (k rows 0..7, m cols 0..39)"); print(d[:40, :8].T.astype(int))
    print(" m = 64..71, 120..127:"); print(d[64:72, :8].T.astype(int)); print(d[120:128, :8].T.astype(int))
