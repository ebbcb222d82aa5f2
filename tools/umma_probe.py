"""GPU probe: issue/commit timing of tcgen05.mma for N = 64/128/256 and several K (mvsn_selftest_umma_probe) --
the per-MMA and per-commit costs quoted in DESIGN.md."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvsnerf_b200 import lib
L = C.CDLL(os.path.join(os.path.dirname(lib.LIB_PATH), 'libmvsnerf_b200_probes.so'))   # python -m mvsnerf_b200.build --probes
L.mvsn_selftest_umma_probe.argtypes = [C.c_void_p]*2 + [C.c_int]*2 + [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
for N in (64, 128, 256):
    for K in (64, 128):
        A = torch.randn(128, K, device="cuda").half(); B = torch.randn(N, K, device="cuda").half()
        D = torch.zeros(128, N, device="cuda"); cyc = torch.zeros(1, dtype=torch.int64, device="cuda")
        reps = 200
        L.mvsn_selftest_umma_probe(lib.ptr(A), lib.ptr(B), N, K, lib.ptr(D), reps, lib.ptr(cyc), None)
        torch.cuda.synchronize()
        n_mma = reps * (K // 64) * 4
        print(f"N={N} K={K}: {cyc.item()} cycles for {n_mma} MMAs -> {cyc.item()/n_mma:.1f} cycles/MMA (ideal {N/2})")
