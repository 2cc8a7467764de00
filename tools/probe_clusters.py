"""Prints how many clusters of the decoder-step projection kernel fit on the device per cluster size, and the splits chosen
for the shapes of a decode position (dev aid for csrc/gemm_x3c.cu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from avsr_b200 import _lib as L  # noqa: E402

lib = L.load()
torch.cuda.init()
torch.zeros(1, device="cuda")
for nb in (32, 96, 128):
    print(f"nb={nb}:", {cs: lib.avsr_dec_proj_max_clusters(cs, nb) for cs in range(1, 17)})
for R in (3, 96, 160, 288):
    print(f"R={R}:", {(N, K): lib.avsr_dec_proj_splits(R, N, K) for N, K in ((3072, 1024), (1024, 1024), (1024, 3072), (5049, 1024))})
