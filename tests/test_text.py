"""avsr_b200/text.py against outputs of the unmodified reference (tests/golden/norm_text.json, oracle/gen_golden_text.py)."""
import json
import os

from avsr_b200 import text as X

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = json.load(open(os.path.join(ROOT, "tests", "golden", "norm_text.json"), encoding="utf8"))


def test_norm_string_matches_reference_outputs():
    bad = [(inp, want, X.norm_string(inp)) for inp, want in G["norm_string"] if X.norm_string(inp) != want]
    assert not bad, bad[:5]
    assert len(G["norm_string"]) > 300


def test_ids_to_text_matches_reference_post_process():
    tl = G["post_process"]["token_list"]
    for ids, want in G["post_process"]["cases"]:
        assert X.ids_to_text(ids, tl) == want, (ids, want)


def test_text_functions_plug_into_the_wer():
    from avsr_b200 import sharding as S
    to_text, norm = X.make_text_functions(["<blank>", "<unk>", "▁it's", "▁a", "▁good", "▁day.", "<eos>"])
    hyp = norm(to_text([2, 3, 4, 5]).replace("<eos>", "").replace("<unk>", ""))
    assert hyp == "IT'S A GOOD DAY"
    assert S.corpus_wer([norm("It's a very good day")], [hyp]) == (1, 5)
