"""Dev aid: step-by-step comparison of the GPU beam search with the CPU oracle (first divergence + margins)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from avsr_b200 import synth
from avsr_b200.weights import DecoderWeights
from avsr_b200.beam_search import BatchedBeamSearch
from oracle import avsr_oracle as O

T, beam = int(sys.argv[1]), int(sys.argv[2])
g = np.load("tests/golden/model_seed0.npz")
sd = synth.make_state_dict(0)
x = torch.from_numpy(g[f"enc_T{T}"])
trace = []
ref = O.beam_search(sd, x, beam, kv_cache=True, trace=trace)
bs = BatchedBeamSearch(DecoderWeights(sd, torch.device("cuda:0")), beam_size=beam, use_graph=False)
nb = bs(x.cuda())
s = bs.last_session
tok = s["hist_tok"].cpu().numpy()[0]; prev = s["hist_prev"].cpu().numpy()[0]
for i, tr in enumerate(trace):
    rt, rp = tr["tok"].tolist(), tr["prev"].tolist()
    gt, gp = tok[i].tolist(), prev[i].tolist()
    w = tr["weighted"].view(-1)
    top = torch.sort(w, descending=True, stable=True)
    vals = top[0][:beam + 3].tolist()
    flag = "" if (rt == gt and rp == gp) else "  <-- DIVERGES"
    print(f"step {i}: ref tok {rt} prev {rp} | gpu tok {gt} prev {gp}{flag}")
    print("   ref top scores:", ["%.6f" % v for v in vals])
    if flag:
        break
print("ref nbest:", [(h.yseq[-4:], round(h.score, 5)) for h in ref])
print("gpu nbest:", [(h.yseq.tolist()[-4:], round(float(h.score), 5)) for h in nb])
print("golden   :", g[f"nbest_T{T}_b{beam}_score"])
print("ref dec/ctc:", [(round(h.dec_score, 5), round(h.ctc_score, 5)) for h in ref])
print("gpu dec/ctc:", [(round(float(h.scores['decoder']), 5), round(float(h.scores['ctc']), 5)) for h in nb])
lp_ref = O.ctc_log_softmax(sd, x.unsqueeze(0))[0]
print("ctc logp max-abs err:", (s["logp"].cpu() - lp_ref).abs().max().item())
kv = O.KVDecoder(sd, x)
ck = torch.stack([k.transpose(0, 1).reshape(-1, 1024) for k in kv.ck])      # [L,T,1024]
cv = torch.stack([v.transpose(0, 1).reshape(-1, 1024) for v in kv.cv])
ckv = s["ckv"].cpu().view(T, 6, 2, 1024)
print("cross K err:", (ckv[:, :, 0].transpose(0, 1) - ck).abs().max().item(), " cross V err:", (ckv[:, :, 1].transpose(0, 1) - cv).abs().max().item())
# last-step decoder log-probs of the running rows (row order = running order at the last step)
dl = s["dec_logp"].cpu()[:beam]
print("last-step dec logp err per row:", [(dl[r] - trace[-1]["dec"][r]).abs().max().item() for r in range(beam)])
# step-0-only run to check single-step numerics
bs2 = BatchedBeamSearch(DecoderWeights(sd, torch.device("cuda:0")), beam_size=beam, use_graph=False)
bs2.decode_batch(x.cuda(), [T], max_steps=1)
s2 = bs2.last_session
print("step0 dec logp err:", (s2["dec_logp"].cpu()[0] - trace[0]["dec"][0]).abs().max().item())
part = s2["part_ids"].cpu()[0].tolist()
psi = s2["psi"].cpu()[0]
print("step0 part ids gpu", part, "ref", trace[0]["part"][0].tolist())
print("step0 ctc psi gpu", psi.tolist(), "ref", [trace[0]["ctc"][0, c].item() for c in part])
