"""GPU probe: error of the 2-term fp16 operand split (hi*hi + hi*lo + lo*hi) through mvsn_selftest_umma, incl.
subnormal lo terms and accumulation order -- the measurements quoted in DESIGN.md for MVSN_MLP_TC_SPLIT."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvsnerf_b200 import lib
L = lib.load()
def st(A, B, N, K):
    D = torch.zeros(128, N, device="cuda")
    lib.check(L.mvsn_selftest_umma(lib.ptr(A.contiguous()), lib.ptr(B.contiguous()), None, N, K, lib.ptr(D), None), "st")
    torch.cuda.synchronize(); return D
g = torch.Generator().manual_seed(0)
A = (torch.rand(128, 128, generator=g) * 16).cuda(); B = (torch.randn(128, 128, generator=g) * 25.6).cuda()
Ah = A.half(); Al = (A - Ah.float()).half(); Bh = B.half(); Bl = (B - Bh.float()).half()
ref = (A.double() @ B.double().t()); mx = ref.abs().max().item()
d_sep = st(Ah, Bh, 128, 128) + st(Ah, Bl, 128, 128) + st(Al, Bh, 128, 128)
d_one = st(torch.cat([Ah, Ah, Al], 1), torch.cat([Bh, Bl, Bh], 1), 128, 384)          # one accumulator, big term first
d_rev = st(torch.cat([Ah, Al, Ah], 1), torch.cat([Bl, Bh, Bh], 1), 128, 384)          # small terms first
print(f"ref max {mx:.1f} | separate accumulators rel err {(d_sep.double()-ref).abs().max().item()/mx:.3e} | "
      f"one accumulator (hi*hi first) {(d_one.double()-ref).abs().max().item()/mx:.3e} | (corrections first) {(d_rev.double()-ref).abs().max().item()/mx:.3e}")
