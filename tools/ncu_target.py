"""ncu target: forwards of one n x h x w tile batch in a given precision, launch by launch (no CUDA graph), so that
`ncu -k regex:conv3x3_umma -s <i> -c <n>` lands on a known conv (launch order = fisr_param_name order; per level:
enc0 0-4, enc1 5-9, enc2 10-14, bottleneck 15-17, dec2 18-23, dec1 24-29, dec0 30-35, FI-SR 36-40, SR 41-45;
levels 1, 2, 3 start at conv 0, 46, 92)."""
import os
import sys

os.environ["FISR_NO_GRAPH"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import fisr_b200  # noqa: E402
from fisr_b200.init import xavier_params  # noqa: E402

n, h, w = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (4, 544, 992)
prec = sys.argv[4] if len(sys.argv) > 4 else "f16f8"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
eng = fisr_b200.Engine(0, precision=prec)
eng.set_params(xavier_params(0, 0.01))
x = torch.rand(n, h, w, 29, generator=torch.Generator().manual_seed(1)).cuda()
for _ in range(reps):
    out = eng.forward(x)
torch.cuda.synchronize()
print("checksum", float(out[2].double().sum()))
