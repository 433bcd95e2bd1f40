"""Bandwidth of the host-side sign/bit-pack (ch_host_pack_sign) over the thread count, on this box's cores.
usage: python dev/host_pack_sweep.py [rows] [nbit]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from concepthash_b200 import _lib as L  # noqa: E402

lib = L.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nbit = int(sys.argv[2]) if len(sys.argv) > 2 else 128
x = np.random.default_rng(0).standard_normal((n, nbit), dtype=np.float32)
out = np.zeros((n, (nbit + 31) // 32), dtype=np.uint32)
fl = C.c_uint32(0)
print("cpus:", os.cpu_count(), "affinity:", len(os.sched_getaffinity(0)))
for nt in (1, 2, 4, 8, 12, 16, 24, 32, 48, 64):
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        lib.ch_host_pack_sign(x.ctypes.data, n, nbit, nbit, out.ctypes.data, C.byref(fl), nt)
        best = min(best, time.perf_counter() - t0)
    print(f"threads {nt:3d}: {best * 1e3:7.2f} ms  {x.nbytes / best / 1e9:7.1f} GB/s")
