"""Debug helper: selftest GEMM with the accumulator at TMEM column MAVA_TC_DCOL."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mava_b200 import native
mode, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
g = torch.Generator().manual_seed(0)
if mode == 0: A, B = torch.randn(128, K, generator=g), torch.randn(K, N, generator=g); ref = A.bfloat16().float() @ B.bfloat16().float()
elif mode == 1: A, B = torch.randn(128, K, generator=g), torch.randn(N, K, generator=g); ref = A.bfloat16().float() @ B.bfloat16().float().T
else: A, B = torch.randn(K, 128, generator=g), torch.randn(K, N, generator=g); ref = A.bfloat16().float().T @ B.bfloat16().float()
D = torch.zeros(128, N, device="cuda")
native.tc_selftest(mode, A.cuda().contiguous(), B.cuda().contiguous(), D, N, K)
torch.cuda.synchronize()
print("dcol", os.environ.get("MAVA_TC_DCOL"), "mode", mode, N, K, "maxerr", float((D.cpu() - ref).abs().max()))
