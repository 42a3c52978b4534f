import sys, os, numpy as np
sys.path.insert(0, "tests")
import test_golden as T
r = T._cuda_render(0, 0)
G = T.G
for k, g in (("ids", "ids_center"), ("ids_s0", "ids_s0")):
    same = np.all(r[k] == G[g], axis=-1)
    print(k, "mismatch", int((~same).sum()), "of", same.size)
    ys, xs = np.nonzero(~same)
    for y, x in list(zip(ys, xs))[:10]:
        print("  px", x, y, "got", r[k][y, x], "want", G[g][y, x], "t", r["tuv"][y, x, 0], G["tuv_center"][y, x, 0])
