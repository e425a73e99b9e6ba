"""Design study (not product): max-abs error of reduced-precision conv operand formats vs fp64.
Decides which MMA operand format meets the north-star 1e-3 max-abs bar."""
import sys, time, torch
sys.path.insert(0, '.')
from oracle import fisrnet_oracle as O

def q_fp16(x): return x.to(torch.float16).to(x.dtype)
def q_bf16(x): return x.to(torch.bfloat16).to(x.dtype)
def q_tf32_trunc(x):
    xi = x.to(torch.float32).view(torch.int32) & ~0x1FFF
    return xi.view(torch.float32).to(x.dtype)
def q_fp16x2(x):
    hi = x.to(torch.float16); lo = (x - hi.to(x.dtype)).to(torch.float16)
    return hi.to(x.dtype) + lo.to(x.dtype)
def q_bf16x2(x):
    hi = x.to(torch.bfloat16); lo = (x - hi.to(x.dtype)).to(torch.bfloat16)
    return hi.to(x.dtype) + lo.to(x.dtype)

H = int(sys.argv[1]) if len(sys.argv) > 1 else 96
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1
for seed in (0, 1):
    p64 = O.init_params(seed, torch.float64)
    p32 = O.cast_params(p64, torch.float32)
    x = O.synthetic_input(N, H, H, seed + 10)
    t = time.time(); ref = O.model(p64, x); t64 = time.time() - t
    t = time.time(); o32 = O.model(p32, x); t32 = time.time() - t
    print(f"seed {seed} H={H} N={N}: fp64 {t64:.2f}s fp32 {t32:.2f}s; out range l3 [{ref[2].min():.3f},{ref[2].max():.3f}] std {ref[2].std():.3f}")
    def rep(tag, outs):
        errs = [ (a.double()-b).abs().max().item() for a, b in zip(outs, ref)]
        ps = [O.psnr(a, b) for a, b in zip(outs, ref)]
        print(f"  {tag:14s} maxabs l1/l2/l3 = {errs[0]:.2e} {errs[1]:.2e} {errs[2]:.2e}   psnr-vs-fp64 l3 {ps[2]:.1f} dB")
    rep('fp32', o32)
    for tag, q in (('fp16', q_fp16), ('tf32-trunc', q_tf32_trunc), ('bf16', q_bf16), ('fp16x2(3mma)', q_fp16x2), ('bf16x2(3mma)', q_bf16x2)):
        rep(tag, O.model(p64, x, operand_hook=q))
