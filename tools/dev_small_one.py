import sys, os
sys.path.insert(0, os.getcwd())
import sparseeigen_b200 as sb
from oracle import speigen_oracle as o
m = int(sys.argv[1]) if len(sys.argv) > 1 else 100
X, Vq = o.spiked_samples(m, -(-m // 5), 3, 0.1, seed=m)
for _ in range(3):
    r = sb.spEigen(X, 3, rho=0.5, data=True)
print(r["info"]["iterations"], r["info"]["backtracks"], r["info"]["init_passes"])
