import sys, os
sys.path.insert(0, "/root/repo")
import torch
import qob200 as Q
from bench import build_chain
n = int(os.environ.get("N", "28"))
B, H = build_chain(Q, n)
x = Q.Ket(B); Q.fill_state(x.data, 7, 2.0 ** (-n / 2)); y = Q.Ket(B)
os.environ.pop("QOB_QREG_STATS", None)
for _ in range(3): Q.mul_(y, H, x, 0.5 - 1j, 0.0)
torch.cuda.synchronize()
os.environ["QOB_QREG_STATS"] = "1"
Q.mul_(y, H, x, 0.5 - 1j, 0.0)
torch.cuda.synchronize()
