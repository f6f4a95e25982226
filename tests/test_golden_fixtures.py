"""CPU checks of the committed full-size fixtures (tests/golden/*.json, written by tests/golden/make_golden.py with the oracle):
they must be discriminating -- 64 distinct teacher inputs, four logits checkpoints per continuation, a diverse greedy path in
the untied variant -- and reproducible: the first layers of the oracle, re-run here, give the recorded embedding-level facts."""
import base64
import json
from pathlib import Path

import numpy as np

from oracle import orc

GOLDEN = Path(__file__).parent / "golden"
VARIANTS = ("bf16", "w4", "bf16-untied")


def load(variant):
    return json.loads((GOLDEN / f"llama1b_L16_{variant}_p512_s64.json").read_text())


def test_fixtures_are_discriminating():
    for variant in VARIANTS:
        g = load(variant)
        tf, gr = g["teacher"], g["greedy"]
        assert g["prompt_len"] == 512 and g["steps"] == 64 and g["config"]["n_layers"] == 16
        assert tf["distinct_inputs"] >= 60, "teacher inputs collapse"
        assert sorted(tf["checkpoints"]) == sorted(gr["checkpoints"]) == ["0", "15", "31", "63"]
        for seq in (tf, gr):
            assert len(seq["inputs"]) == len(seq["argmax_after"]) == len(seq["second_after"]) == len(seq["top2_gap_ulps"]) == 64
            for cp in seq["checkpoints"].values():
                sub = np.frombuffer(base64.b64decode(cp["every16_b64"]), dtype=np.uint16)
                assert len(sub) == 128256 // 16 and len(cp["top8_ids"]) == 8
                top = orc.bf16_to_f32(np.array(cp["top8_bits"], np.uint16))
                assert np.all(np.diff(top) <= 0), "top-8 not in descending order"
        # the greedy continuation feeds its own argmax back
        assert gr["inputs"][0] == g["prompt_argmax"]
        assert gr["inputs"][1:] == gr["argmax_after"][:-1]
    assert load("bf16-untied")["greedy"]["distinct_inputs"] >= 32, "the untied variant must decode a diverse greedy path"


def test_fixture_inputs_are_the_documented_hash_streams():
    vocab = 128256
    want = [int(orc.lib().orc_hash_int(0x5EED, 0xFFFE, i, 0, vocab)) for i in range(64)]
    for variant in VARIANTS:
        assert load(variant)["teacher"]["inputs"] == want
