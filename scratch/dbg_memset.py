import ctypes, os, torch
print("alloc conf", os.environ.get("PYTORCH_CUDA_ALLOC_CONF"), os.environ.get("PYTORCH_ALLOC_CONF"), torch.cuda.get_allocator_backend())
rt = ctypes.CDLL("libcudart.so.12")
st = ctypes.c_void_p()
print("create", rt.cudaStreamCreate(ctypes.byref(st)))
for nbytes, ms in ((1 << 20, 7360), (4 << 20, 28800), (4 << 20, 65280), (9 << 20, 65280)):
    t = torch.empty(0, dtype=torch.uint8, device="cuda")
    t.resize_(nbytes)
    p = t.data_ptr()
    rc = rt.cudaMemsetAsync(ctypes.c_void_p(p), 0, ctypes.c_size_t(ms), st)
    rc0 = rt.cudaMemsetAsync(ctypes.c_void_p(p), 0, ctypes.c_size_t(ms), ctypes.c_void_p(0))
    print(nbytes, ms, hex(p), "memset on created stream:", rc, " on default stream:", rc0, "sync", rt.cudaDeviceSynchronize())
