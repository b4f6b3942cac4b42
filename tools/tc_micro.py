import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vtaco_b200 import build as _b
torch.zeros(1, device='cuda')
L = C.CDLL(_b.BENCH_LIB)   # debug kernels are not part of the product library
L.vtaco_tc_microbench.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
out = (C.c_longlong * 6)()
for N, n_acc in ((32, 1), (32, 4), (64, 1), (256, 1)):
    for rounds in (1, 2, 8):
        L.vtaco_tc_microbench(rounds, n_acc, N, out)
        n = 12 * rounds
        print('N=%3d acc=%d n_mma=%2d : issue %5d cyc, issue+complete %5d cyc  (%.1f cyc/MMA)' % (N, n_acc, n, out[4], out[5], out[5] / n))
