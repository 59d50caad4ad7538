import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lb.lib().lb200_set_batched_mode(mode)
a = lb.dev.larnv_matrix(32, 32 * batch).t().contiguous().view(batch, 32, 32)
for _ in range(2):
    lb.dev.getrf_batched32(a)
torch.cuda.synchronize()
