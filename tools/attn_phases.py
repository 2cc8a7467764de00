"""Dev aid: phase timestamps of the self-attention kernel inside a real decode (CTA (0,0), last launch before each read)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avsr_b200 import _lib as L, synth
from avsr_b200.model import AVSRCocktailB200
pos = int(sys.argv[1]) if len(sys.argv) > 1 else 187
m = AVSRCocktailB200(synth.make_state_dict(0), beam_size=3)
bs = m.beam_search
x = torch.nn.functional.layer_norm(torch.randn(32 * 375, 1024, device="cuda"), (1024,))
buf = torch.zeros(8, dtype=torch.int64, device="cuda")
lib = L.load()
bs.decode_batch(x, [375] * 32, max_steps=pos)
lib.avsr_dec_attn_debug(L.ptr(buf))
s = bs.last_session
names = ["entry", "after wait", "query gather", "row list start", "row list done", "tile loop start", "tile loop done", "output stored"]
g = torch.cuda.CUDAGraph()
bs._skip = frozenset({"advance"})
bs._step(s); torch.cuda.synchronize()
with torch.cuda.graph(g):
    for _ in range(4):
        bs._step(s)
for rep in range(3):
    g.replay(); torch.cuda.synchronize()
    t = buf.cpu().tolist()
    print(f"position {int(s['step'].item())}: " + "  ".join(f"{names[i]} +{(t[i] - t[i-1]) / 1e3:.2f}us" for i in range(1, 8)) + f"   total {(t[7]-t[0])/1e3:.2f}us")
lib.avsr_dec_attn_debug(None)
