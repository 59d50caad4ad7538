"""C1 diagnostics: solution differences at n=4096 between the oracle, the two few-RHS solve modes and XACT."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
from oracle import oracle as O
SEED = (1988, 1989, 1990, 1991)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
a, seed = O.random_matrix(n, n, SEED)
xact, _ = O.random_matrix(n, 1, seed)
b = np.asfortranarray(a @ xact)
lu_ref, x_ref = a.copy(order="F"), b.copy(order="F")
ipiv_ref, info_ref = O.dgesv(lu_ref, x_ref)
rel = lambda x, y: float(np.max(np.abs(x - y)) / np.max(np.abs(y)))
print("oracle vs xact", rel(x_ref, xact))
lu = a.copy(order="F"); ipiv = np.zeros(n, dtype=np.int32)
lb.f77.dgetrf(n, n, lu, n, ipiv)
print("ipiv equal", np.array_equal(ipiv, ipiv_ref), "lu rel", rel(lu, lu_ref))
for mode in (1, 0):
    lb.lib().lb200_set_fewrhs_mode(mode)
    x = b.copy(order="F"); lb.f77.getrs("N", lu, ipiv, x)
    print(f"mode {mode}: gpu-LU solve vs oracle {rel(x, x_ref):.3e}  vs xact {rel(x, xact):.3e}")
    x2 = b.copy(order="F"); lb.f77.getrs("N", lu_ref, ipiv_ref, x2)
    print(f"mode {mode}: oracle-LU solve vs oracle {rel(x2, x_ref):.3e}  vs xact {rel(x2, xact):.3e}")
    x3 = b.copy(order="F"); O.dgetrs("N", lu, ipiv, x3)
    print(f"         oracle solve on gpu-LU vs oracle {rel(x3, x_ref):.3e}; gpu solve vs oracle solve on same LU {rel(x, x3):.3e}")
