import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
print(bench.block_microbench(torch.device("cuda:0"), 256, iters=1, profile=True))
