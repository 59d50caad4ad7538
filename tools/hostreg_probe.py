"""How fast is cudaHostRegister on this box, and does it parallelise over threads?  python tools/hostreg_probe.py [GiB]"""
import ctypes, glob, os, sys, threading, time
import numpy as np, torch
torch.cuda.init(); torch.zeros(1, device="cuda")
path = [p for p in glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + glob.glob("/usr/local/cuda/lib64/libcudart.so*")][0]
rt = ctypes.CDLL(path)
rt.cudaHostRegister.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint]
rt.cudaHostUnregister.argtypes = [ctypes.c_void_p]
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
nbytes = int(gib * (1 << 30))
buf = np.empty(nbytes, dtype=np.uint8); buf[::4096] = 1      # touch every page
base = buf.ctypes.data
for nth in (1, 2, 4, 8, 16):
    part = (nbytes // nth) & ~0xFFFFF
    res = [None] * nth
    def work(i):
        torch.cuda.set_device(0)
        length = part if i < nth - 1 else nbytes - part * (nth - 1)
        res[i] = rt.cudaHostRegister(base + i * part, length, 0)
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i,)) for i in range(nth)]
    [t.start() for t in th]; [t.join() for t in th]
    t1 = time.perf_counter()
    th = [threading.Thread(target=lambda i=i: rt.cudaHostUnregister(base + i * part)) for i in range(nth)]
    [t.start() for t in th]; [t.join() for t in th]
    t2 = time.perf_counter()
    print(f"{gib} GiB, {nth:2d} threads: register {1e3 * (t1 - t0):7.1f} ms, unregister {1e3 * (t2 - t1):7.1f} ms, rc={set(res)}", flush=True)
