"""Bring-up probe of the tensor-memory A operand (csrc/probes/tc_ts_probe.cu): per round
D = X[M,64] B1^T + H[M,128] B2^T with X from shared memory and H from TMEM (tcgen05.st by the row's thread),
single CTA (M=128) and CTA pair (M=256, B split, remote mbarrier arrive, multicast commit).

    python -m mvsnerf_b200.build --probes        # here (nvcc cross-compiles)
    gpurun -- 'timeout 60 python tools/ts_probe.py'
"""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L = C.CDLL(os.path.join(ROOT, "mvsnerf_b200", "libmvsnerf_b200_probes.so"))
L.mvsn_probe_umma_ts.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
L.mvsn_probe_umma_ts.restype = C.c_int
L.mvsn_probe_last_error.restype = C.c_char_p
ok = True
for pair in (0, 1):
    M, rounds = 128 * (1 + pair), 5
    g = torch.Generator(device="cuda").manual_seed(7 + pair)
    X = (torch.randn(rounds, M, 64, device="cuda", generator=g) * 0.5).half()
    H = (torch.randn(rounds, M, 128, device="cuda", generator=g) * 0.5).half()
    B1 = (torch.randn(128, 64, device="cuda", generator=g) * 0.5).half()
    B2 = (torch.randn(128, 128, device="cuda", generator=g) * 0.5).half()
    D = torch.full((rounds, M, 128), float("nan"), device="cuda")
    rc = L.mvsn_probe_umma_ts(X.data_ptr(), H.data_ptr(), B1.data_ptr(), B2.data_ptr(), rounds, pair, D.data_ptr(), None)
    assert rc == 0, L.mvsn_probe_last_error()
    try:
        torch.cuda.synchronize()
    except Exception as e:          # a trapped mbarrier wait = a broken hand-off
        print(f"pair={pair}: kernel failed: {e}")
        sys.exit(1)
    ref = X.float() @ B1.float().t() + H.float() @ B2.float().t()
    err = (D - ref).abs().amax(dim=(1, 2))
    only_x = (D - X.float() @ B1.float().t()).abs().max().item()
    print(f"pair={pair} M={M}: per-round max |D - ref| = {[f'{e:.2e}' for e in err.tolist()]}   (|D - X B1^T| = {only_x:.2e})")
    ok &= bool((err < 2e-3 * ref.abs().max()).all())
print("TS PROBE", "PASS" if ok else "FAIL")
sys.exit(0 if ok else 1)
