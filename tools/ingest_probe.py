"""Host-ingest probe of a multi-GPU box (one process, all GPUs): every GPU copies `gib` GiB device -> pinned host in
64 MiB pieces at the same time.  Compares ordinary pinned memory (cudaHostAlloc, 4 KiB pages) with a transparent-huge-page
region registered with cudaHostRegister (2 MiB pages: far fewer IOMMU translations per byte -- the suspect for a VM whose
eight GPUs together only reach ~94 GB/s), and with write-combined pinned memory.

    python tools/ingest_probe.py [gib_per_gpu]
"""
import ctypes
import mmap
import sys
import time

import numpy as np
import torch

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
n_dev = torch.cuda.device_count()
nbytes = int(gib * 2**30)
rt = torch.cuda.cudart()
libc = ctypes.CDLL("libc.so.6", use_errno=True)
MADV_HUGEPAGE = 14


def thp_buffer():
    m = mmap.mmap(-1, nbytes + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    addr = ctypes.addressof(ctypes.c_char.from_buffer(m))
    al = (addr + (2 << 20) - 1) & ~((2 << 20) - 1)
    rc = libc.madvise(ctypes.c_void_p(al), ctypes.c_size_t(nbytes), MADV_HUGEPAGE)
    a = np.frombuffer(m, dtype=np.uint8, count=nbytes, offset=al - addr)
    a[:: 4096] = 1                                    # fault the pages in
    err = rt.cudaHostRegister(al, nbytes, 0)
    t = torch.from_numpy(a)
    return t, m, rc, int(err)


def run(label, host):
    devs = [torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{d}") for d in range(n_dev)]
    streams = [torch.cuda.Stream(device=d) for d in range(n_dev)]
    for subset in ([0], list(range(min(2, n_dev))), list(range(min(4, n_dev))), list(range(n_dev))):
        if len(subset) > n_dev:
            continue
        for d in range(n_dev):
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        for d in subset:
            with torch.cuda.device(d), torch.cuda.stream(streams[d]):
                for a, b in zip(host[d].split(64 << 20), devs[d].split(64 << 20)):
                    a.copy_(b, non_blocking=True)
        for d in subset:
            streams[d].synchronize()
        el = time.perf_counter() - t0
        print(f"{label}: {len(subset)} GPU(s) {len(subset) * nbytes / el / 1e9:.1f} GB/s aggregate", flush=True)


print("THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip())
pinned = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(n_dev)]
run("cudaHostAlloc (4 KiB pages)", pinned)
del pinned
keep = [thp_buffer() for _ in range(n_dev)]
print("madvise rc", [k[2] for k in keep], "cudaHostRegister rc", [k[3] for k in keep])
try:
    print("AnonHugePages:", [l for l in open("/proc/meminfo") if "AnonHugePages" in l][0].strip())
except Exception:
    pass
run("THP + cudaHostRegister (2 MiB pages)", [k[0] for k in keep])
