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


def test_chunking_and_timestamps_match_reference_outputs():
    S = json.load(open(os.path.join(ROOT, "tests", "golden", "stitch.json")))
    for c in S["chunks"]:
        assert [list(s) for s in X.fixed_chunks(c["duration"], c["max_length"])] == c["segments"], c
    for t, want in S["timestamps"]:
        assert X.format_vtt_timestamp(t) == want, t
    import pytest
    with pytest.raises(ValueError):
        X.fixed_chunks(0.0)


def test_vtt_and_stitching():
    hyps = X.segment_hypotheses([(0.0, 4.0), (4.0, 8.5), (8.5, 9.0)], [" HELLO <unk> WORLD ", "<unk>", "IT'S 5%"], offset=3600.25)
    assert hyps[1] == {"start_time": 3604.25, "end_time": 3608.75, "text": "<unk>"}
    assert X.write_vtt(hyps) == ("WEBVTT\n\n01:00:00.250 --> 01:00:04.250\nHELLO  WORLD\n\n"
                                 "01:00:08.750 --> 01:00:09.250\nIT'S 5%\n\n")
    assert X.write_vtt([]) == "WEBVTT\n\n"
    # chunk outputs arrive in dataset order; the reference sorts them by (start time, text) before the WER
    assert X.stitch_outputs([8.0, 0.0, 4.0, 4.0], ["the end.", "Well-known <unk>", "b", "a"]) == "WELL KNOWN A B THE END"
    import pytest
    with pytest.raises(ValueError):
        X.stitch_outputs([0.0], [])
