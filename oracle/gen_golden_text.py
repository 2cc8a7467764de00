"""Generates tests/golden/norm_text.json: input / output pairs of the reference's text normalisation and id -> text
post-processing, produced by the UNMODIFIED reference (/root/reference/src/tokenizer/norm_text.py:121-134 `norm_string`,
src/tokenizer/spm_tokenizer.py:46-54 `TextTransform.post_process`).  Run in the build container only (the reference is not
present on the GPU box); the committed JSON is what tests/test_text.py checks avsr_b200/text.py against."""
import json
import os
import random
import sys

sys.path.insert(0, "/root/reference")
from src.tokenizer.norm_text import norm_string          # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cases = [
    "I'm Binh i'm 25 years old i'm a AI researcher. It's a good day.", "t_qua ng'_123", "", "   ", "hello", "Hello, world!", "well-known fact",
    "state-of-the-art systems", "50% of people", "12.5% growth", "it costs $5", "it costs $1,200.50 today", "5$ each", "£30 only", "1,000£ fine",
    "pi is 3.14", "1,5 or 2,75", "at 5 p.m. sharp", "the U.S.A. team", "visit example.com now", "www.site.co.uk/path", "don't can't won't it's we're they've i'll",
    "rock'n'roll", "O'Neil", "’tis the season", "she said \"hello\"", "(parenthetical) [bracketed] {braced}", "semi;colon: colon", "a/b c\\d", "e=mc^2",
    "hash #tag @user", "under_score", "tilde~ back`tick", "what?! really...", "-dash- --double--", "word<|en|><|transcribe|>", "Hi!<|ru|><|translate|>",
    "100", "3.0", "v1.2.3", "a.b", "A.B.C", "x-ray y-axis", "co-op's", "naïve café", "ÀÉÎ", "mixed123abc", "123abc456", "$", "%", "£", ".", "'", "' '", "a  b   c",
    "tab\tseparated\nnewline", "end.", "end,", "end!", "end?", "end;", "end:", "50%.", "$5.", "(50%)", "\"$5\"", "1,000,000", "1.000.000", "3,14%", "0.5$", "$0.5",
    "i.e.", "e.g.", "p.m", "a.m.", "U.S", "foo.bar.baz", "foo..bar", "x+y", "+1", "a|b", "<unk>", "hello <unk> world", "it’s", "rock’n’roll",
]
rng = random.Random(7)
alphabet = "abcXYZ019 .,!?;:'\"-][~+%$£<>/()_=@#^\\`{}|’"
for _ in range(300):
    cases.append("".join(rng.choice(alphabet) for _ in range(rng.randint(1, 14))))
out = {"norm_string": [[c, norm_string(c)] for c in cases]}

# id -> text (post_process) on a tiny synthetic token list in the same format as the reference's unit file
tok = ["<blank>", "<unk>", "▁THE", "▁CAT", "S", "▁SAT", "'", "T", "<space>", "▁", "<eos>"]
import torch                                           # noqa: E402


def post_process(token_ids):                           # the reference method, bound to the synthetic list
    from src.tokenizer.spm_tokenizer import TextTransform
    obj = TextTransform.__new__(TextTransform)
    obj.token_list = tok
    return TextTransform.post_process(obj, torch.tensor(token_ids))


seqs = [[2, 3, 4, 5], [2, -1, 3, -1], [1, 2, 8, 3], [9, 2, 9], [], [5, 6, 7, 10], [3, 10]]
out["post_process"] = {"token_list": tok, "cases": [[s, post_process(s)] for s in seqs]}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "norm_text.json"), "w"), ensure_ascii=False, indent=0)
print(len(cases), "norm_string cases,", len(seqs), "post_process cases")
