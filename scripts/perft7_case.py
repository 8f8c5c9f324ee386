import sys, numpy as np
sys.path.insert(0, "quantik-core-py_b200")
import quantik_b200 as Q
e = np.zeros((1, 8), np.uint16)
for _ in range(2):
    p = Q.perft(e, 7)
print(int(p["nodes"][7]))
