import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
r = bench.cnn_bar(bench.CONFIGS["c3"], torch.device("cuda"), 9)
print(r["ms"])
