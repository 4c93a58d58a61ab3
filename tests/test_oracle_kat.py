"""Pin both CPU oracles to the known-answer vectors of SURVEY.md §A.8 (tests/golden/kat_survey.json).

The reference ships no golden vectors ("parity unpinned" upstream); these KATs come from a third,
independent model, so agreement here means three restatements of the VHDL agree."""
import json
import os

import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import py_oracle as po

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat_survey.json")))


def _g(d):
    base = dict(nfft_log2=3, data_width=16, twdl_width=16, format=0, rndmode=0, xser=1, use_fly=1, direction=0)
    base.update(d)
    return base


@pytest.mark.parametrize("case", KAT["twiddles"], ids=lambda c: f"s{c['stage']}_x{c['xser']}_tw{c['tw']}")
def test_twiddle_kat(case):
    g = co.generics(12, twdl_width=case["tw"], xser=case["xser"])
    re, im = co.twiddle_table(g, case["stage"])
    for k, (wr, wi) in zip(case["k"], case["w"]):
        assert (int(re[k]), int(im[k])) == (wr, wi), f"C oracle k={k}"
        assert po.twiddle(case["stage"], k, case["tw"], case["xser"]) == (wr, wi), f"py oracle k={k}"


@pytest.mark.parametrize("case", KAT["frames"], ids=lambda c: c["name"])
def test_frame_kat(case):
    gd = _g(case["generics"])
    x = np.array(case["in"], np.int64)
    want = [tuple(v) for v in case["out"]]
    ore, oim = co.transform(co.generics(**gd), x[:, 0], x[:, 1])
    assert list(zip(ore.tolist(), oim.tolist())) == want
    assert po.transform(po.Generics(**gd), [tuple(v) for v in case["in"]]) == want
    # and through the batched/container path
    out = co.batch(co.generics(**gd), x.astype(co.scalar_dtype(gd["data_width"])).reshape(1, -1, 2))
    assert [tuple(v) for v in out[0].tolist()] == want
