"""Round-2 bring-up probe for `cta_group::2` (csrc/tc_pair_probe.cu): one M=256 x N=128 x K GEMM by a CTA pair with
the B operand split across the two CTAs.  Run on a B200 under a timeout:

    timeout 60 python tools/umma_pair_probe.py

Prints the max error against torch (B200, end of round 1: 4.8e-6 / 6.7e-6 / 1.5e-5 for K = 64 / 128 / 256)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvsnerf_b200 import lib
L = C.CDLL(os.path.join(os.path.dirname(lib.LIB_PATH), 'libmvsnerf_b200_probes.so'))   # python -m mvsnerf_b200.build --probes
L.mvsn_probe_umma_pair.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
L.mvsn_probe_umma_pair.restype = C.c_int
for K in (64, 128, 256):
    g = torch.Generator(device="cuda").manual_seed(K)
    A = (torch.randn(256, K, device="cuda", generator=g) * 0.5).half()
    B = (torch.randn(128, K, device="cuda", generator=g) * 0.5).half()
    D = torch.full((256, 128), float("nan"), device="cuda")
    rc = L.mvsn_probe_umma_pair(A.data_ptr(), B.data_ptr(), K, D.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    print(f"K={K}: max |D - A B^T| = {(D - ref).abs().max().item():.3e}  (rows 0-127 {(D[:128] - ref[:128]).abs().max().item():.3e}, "
          f"rows 128-255 {(D[128:] - ref[128:]).abs().max().item():.3e})")
