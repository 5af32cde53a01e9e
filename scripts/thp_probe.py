import ctypes, numpy as np, time, re
libc = ctypes.CDLL("libc.so.6", use_errno=True)
def anon_huge():
    s = open("/proc/self/smaps_rollup").read()
    return int(re.search(r"AnonHugePages:\s+(\d+)", s).group(1))
def run(advise, populate):
    a = np.empty(20_000_000, np.int32)
    addr = a.ctypes.data; n = a.nbytes
    lo = (addr + (2 << 20) - 1) & ~((2 << 20) - 1); hi = (addr + n) & ~((2 << 20) - 1)
    rc = 0
    if advise:
        rc = libc.madvise(ctypes.c_void_p(lo), ctypes.c_size_t(hi - lo), 14)
    t = time.perf_counter()
    if populate:
        plo = addr & ~4095; phi = (addr + n + 4095) & ~4095
        rc2 = libc.madvise(ctypes.c_void_p(plo), ctypes.c_size_t(phi - plo), 23)
    else:
        a[::1024] = 1; rc2 = 0
    dt = time.perf_counter() - t
    print(f"advise={advise} populate={populate} madvise_rc={rc} pop_rc={rc2} errno={ctypes.get_errno()} fault_ms={1e3*dt:.2f} AnonHugePages_kB={anon_huge()}")
    del a
for adv in (0, 1):
    for pop in (0, 1):
        run(adv, pop); run(adv, pop)
