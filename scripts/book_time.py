import sys, time, json
sys.path.insert(0, "quantik-core-py_b200")
import quantik_b200 as Q, torch
Q.opening_book_summary(16)
torch.cuda.synchronize(); t0 = time.perf_counter()
s = Q.opening_book_summary(16)
torch.cuda.synchronize(); print("book(16) seconds", time.perf_counter() - t0, s["totals"] if "totals" in s else list(s)[:5])
