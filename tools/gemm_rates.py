"""Times the dense FP64 tensor-core GEMM on resident random operands:  python tools/gemm_rates.py M N K [symmetric] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparseeigen_b200 as sb    # noqa: E402

M, N, K = (int(x) for x in sys.argv[1:4])
sym = bool(int(sys.argv[4])) if len(sys.argv) > 4 else False
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
ms = sb.dense_gemm_time(M, N, K, True, True, sym, reps=reps)
print(list(ms), "TFLOP/s (full product) %.2f" % (2.0 * M * N * K / min(ms) / 1e9))
