import sys, time, torch
sys.path.insert(0, ".")
from concepthash_b200 import hashing
ev = hashing.get_evaluator()
b = ev.b._b
b.begin()
n = 1_000_000
ids = torch.randint(101, (n,))
for name, t in [("pageable i64", ids), ("pinned i64", ids.pin_memory()), ("pageable i32", ids.to(torch.int32)), ("cuda i64", ids.cuda())]:
    for rep in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = b.pack_labels(t, 0xffffffff)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        if rep >= 3:
            print(name, "host ms %.3f  total ms %.3f" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
# raw copies for comparison
pin = torch.empty(n, dtype=torch.int64).pin_memory()
dev = torch.empty(n, dtype=torch.int64, device="cuda")
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter(); pin.copy_(ids); t1 = time.perf_counter()
    dev.copy_(pin, non_blocking=True); torch.cuda.synchronize(); t2 = time.perf_counter()
    dev.copy_(ids); torch.cuda.synchronize(); t3 = time.perf_counter()
    print("torch: host->pinned memcpy %.3f  pinned->dev %.3f  pageable->dev %.3f" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3))
