"""GPU box: can the page cache be the DMA source?  mmap a wav-sized file (tmpfs and the box's disk), cudaHostRegister
the mapping (read-only flag), copy to the device asynchronously, compare with pread into pinned memory.

    python tools/hostreg_probe.py
"""
import ctypes, mmap, os, sys, tempfile, time
import numpy as np
import torch

rt = ctypes.CDLL("libcudart.so.12")
rt.cudaHostRegister.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint]
rt.cudaHostUnregister.argtypes = [ctypes.c_void_p]
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
N, SZ = 256, 960044
torch.cuda.init(); torch.zeros(1, device="cuda")
dev = torch.empty(N * SZ, dtype=torch.uint8, device="cuda")
pin = torch.empty(N * SZ, dtype=torch.uint8).pin_memory()
for where in ("/dev/shm", tempfile.gettempdir()):
    d = tempfile.mkdtemp(dir=where)
    data = np.random.default_rng(0).integers(0, 255, SZ, dtype=np.uint8)
    paths = []
    for i in range(N):
        p = os.path.join(d, "f%04d.bin" % i); data.tofile(p); paths.append(p)
    # a) pread into pinned, one H2D
    t0 = time.perf_counter()
    pv = pin.numpy()
    for i, p in enumerate(paths):
        fd = os.open(p, os.O_RDONLY); os.preadv(fd, [memoryview(pv[i * SZ:(i + 1) * SZ])], 0); os.close(fd)
    t_read = time.perf_counter() - t0
    dev.copy_(pin, non_blocking=True); torch.cuda.synchronize()
    t_a = time.perf_counter() - t0
    # b) mmap + register + async copy + unregister
    for flags, name in ((0x08, "read-only"), (0x00, "default")):
        ok, t_reg, maps = 0, 0.0, []
        t0 = time.perf_counter()
        for i, p in enumerate(paths):
            fd = os.open(p, os.O_RDONLY)
            m = mmap.mmap(fd, SZ, flags=mmap.MAP_SHARED | getattr(mmap, "MAP_POPULATE", 0), prot=mmap.PROT_READ); os.close(fd)
            addr = ctypes.addressof(ctypes.c_char.from_buffer_copy(b"")) if False else None
            buf = (ctypes.c_char * SZ).from_buffer_copy(b"") if False else None
            a = np.frombuffer(m, dtype=np.uint8)
            ptr = a.ctypes.data
            t1 = time.perf_counter()
            rc = rt.cudaHostRegister(ptr, SZ, flags)
            t_reg += time.perf_counter() - t1
            if rc == 0:
                ok += 1
                rt.cudaMemcpyAsync(dev.data_ptr() + i * SZ, ptr, SZ, 1, None)
            maps.append((m, a, ptr, rc))
        torch.cuda.synchronize()
        t_copy = time.perf_counter() - t0
        for m, a, ptr, rc in maps:
            if rc == 0: rt.cudaHostUnregister(ptr)
        t_b = time.perf_counter() - t0
        good = bool((dev[:SZ].cpu().numpy() == data).all()) if ok else None
        print("%-9s %-9s: registered %d/%d (rc of the last %d)  register %.1f us/file  mmap+register+H2D %.3f s (%.1f GB/s)  +unregister %.3f s  data ok %s | pread->pinned %.3f s (%.1f GB/s) + H2D = %.3f s"
              % (where, name, ok, N, maps[-1][3], 1e6 * t_reg / N, t_copy, N * SZ / t_copy / 1e9, t_b, good, t_read, N * SZ / t_read / 1e9, t_a), flush=True)
        del maps
    for p in paths: os.remove(p)
    os.rmdir(d)
