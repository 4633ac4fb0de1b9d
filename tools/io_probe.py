import time, os, ctypes as C, numpy as np, torch
torch.cuda.init(); torch.zeros(1, device="cuda")
rt = C.CDLL("libcudart.so.12")
for mib in (64, 256, 1024):
    p = C.c_void_p()
    t = time.perf_counter(); rt.cudaHostAlloc(C.byref(p), C.c_size_t(mib << 20), 0); dt = time.perf_counter() - t
    t = time.perf_counter(); rt.cudaFreeHost(p); df = time.perf_counter() - t
    print(f"cudaHostAlloc {mib} MiB: {dt*1e3:.1f} ms, free {df*1e3:.1f} ms")
n = 1 << 30
a = np.random.default_rng(0).integers(0, 256, n, dtype=np.uint8)
t = time.perf_counter(); a.tofile("/dev/shm/x.bin"); print(f"write 1 GiB tmpfs: {time.perf_counter()-t:.3f} s")
buf = np.empty(n, np.uint8)
t = time.perf_counter()
with open("/dev/shm/x.bin", "rb") as f: f.readinto(buf)
print(f"read 1 GiB tmpfs: {time.perf_counter()-t:.3f} s")
m = np.empty(n, np.uint8)
t = time.perf_counter(); m[:] = 1; print(f"first touch 1 GiB: {time.perf_counter()-t:.3f} s")
t = time.perf_counter(); rt.cudaHostRegister(C.c_void_p(m.ctypes.data), C.c_size_t(n), 0); print(f"cudaHostRegister 1 GiB (touched): {time.perf_counter()-t:.3f} s")
d = torch.empty(n, dtype=torch.uint8, device="cuda")
pg = torch.from_numpy(buf)
torch.cuda.synchronize(); t = time.perf_counter(); d.copy_(pg); torch.cuda.synchronize(); print(f"H2D 1 GiB pageable: {time.perf_counter()-t:.3f} s")
torch.cuda.synchronize(); t = time.perf_counter(); pg.copy_(d); torch.cuda.synchronize(); print(f"D2H 1 GiB pageable: {time.perf_counter()-t:.3f} s")
os.unlink("/dev/shm/x.bin")
