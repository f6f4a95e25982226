"""Probe (run on the GPU box): can a read-only file mapping be page-locked in place (cudaHostRegister), and what does it buy for H2D copies?
Used to decide how metalchat_b200/csrc/mc_loader.cu feeds the device (loader)."""
import ctypes as C
import mmap
import os
import tempfile
import time

import torch

rt = C.CDLL("libcudart.so.12")
torch.cuda.init()
torch.zeros(1, device="cuda")
n = 256 << 20
path = os.path.join(tempfile.gettempdir(), "hostreg_probe.bin")
with open(path, "wb") as f:
    f.write(os.urandom(1 << 20) * (n >> 20))
attr = C.c_int(-1)
rt.cudaDeviceGetAttribute(C.byref(attr), 113, 0)  # cudaDevAttrHostRegisterReadOnlySupported (driver_types.h)
print("cudaDevAttrHostRegisterReadOnlySupported:", attr.value)
dst = torch.empty(n, dtype=torch.uint8, device="cuda")
libc = C.CDLL(None, use_errno=True)
libc.mmap.restype = C.c_void_p
libc.mmap.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_long]
rt.cudaHostRegister.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]
rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]


def h2d(ptr, what):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rc = rt.cudaMemcpy(C.c_void_p(dst.data_ptr()), C.c_void_p(ptr), n, 1)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"  {what}: rc {rc}, {n / dt / 1e9:.2f} GB/s")


fd = os.open(path, os.O_RDONLY)
for prot, flags, reg, name in [(1, 2, 8, "PROT_READ MAP_PRIVATE + cudaHostRegisterReadOnly"), (1, 2, 0, "PROT_READ MAP_PRIVATE + default flags"),
                               (3, 2, 0, "PROT_READ|WRITE MAP_PRIVATE + default flags"), (1, 1, 8, "PROT_READ MAP_SHARED + ReadOnly")]:
    p = libc.mmap(None, n, prot, flags, fd, 0)
    print(name)
    h2d(p, "pageable (first touch)")
    h2d(p, "pageable (page cache warm)")
    t0 = time.perf_counter()
    rc = rt.cudaHostRegister(C.c_void_p(p), n, reg)
    print(f"  cudaHostRegister rc {rc} in {time.perf_counter() - t0:.3f} s")
    rt.cudaGetLastError()
    if rc == 0:
        h2d(p, "registered")
        rt.cudaHostUnregister(C.c_void_p(p))
    libc.munmap(C.c_void_p(p), n)
pin = torch.empty(n, dtype=torch.uint8).pin_memory()
h2d(pin.data_ptr(), "cudaHostAlloc'ed buffer (reference point)")
os.close(fd)
os.remove(path)
