"""How long do CUDA context creation and the workspace-pool allocation take? (scratch probe)"""
import ctypes, time, sys
t0 = time.perf_counter()
L = ctypes.CDLL("longcalld_b200/csrc/liblcd_gpu.so")
L.lcd_gpu_init.argtypes = [ctypes.c_int, ctypes.c_size_t]
cu = ctypes.CDLL("libcudart.so.12") if False else None
gb = float(sys.argv[1])
t1 = time.perf_counter()
rc = L.lcd_gpu_init(0, int(gb * (1 << 30)))
t2 = time.perf_counter()
L.lcd_gpu_shutdown()
t3 = time.perf_counter()
print(f"pool {gb} GiB: dlopen {t1-t0:.3f} s, lcd_gpu_init {t2-t1:.3f} s (rc {rc}), shutdown {t3-t2:.3f} s")
