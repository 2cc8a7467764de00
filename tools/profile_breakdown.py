"""Dev aid: where does one bench step go?  CUDA-event timings of the encoder stages and the decode phases."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avsr_b200 import synth, _lib as L
from avsr_b200.model import AVSRCocktailB200

B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 32, int(sys.argv[2]) if len(sys.argv) > 2 else 375
sd = synth.make_state_dict(0)
m = AVSRCocktailB200(sd, beam_size=3)
v, a = synth.make_inputs(1234, T, B)
v, a = v.cuda(), a.cuda()

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

enc = m.encoder
vp = v.reshape(B * T, 88, 88).contiguous()
for it in range(2):
    t0 = ev(); x = enc.forward_packed(vp, a, [T] * B); t1 = ev()
    bs = m.beam_search
    bs.n_groups = 1                   # one chain: this tool looks at a single session
    s = bs._session(B, T, B * T); bs.last_session = s
    t2 = ev(); bs.prepare(s, x, [T] * B); t3 = ev()
    nb = bs.decode_batch(x, [T] * B); t4 = ev()
    torch.cuda.synchronize()
    print(json.dumps({"encoder_ms": t0.elapsed_time(t1), "decode_prepare_ms": t2.elapsed_time(t3), "decode_total_ms(incl prepare)": t3.elapsed_time(t4)}))
# encoder stage split: frontend+trunk vs rest
torch.cuda.synchronize()
ft = torch.arange(T, dtype=torch.int32).repeat(B).cuda(); fT = torch.full((B * T,), T, dtype=torch.int32).cuda()
t0 = ev(); enc._video_frontend(vp, ft, fT, B * T); t1 = ev(); torch.cuda.synchronize()
print(json.dumps({"video_frontend_ms": t0.elapsed_time(t1)}))
