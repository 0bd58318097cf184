"""The oracle's restatement of WanModel.forward (oracle.dit_forward; reference model.py:410-497) against the golden
outputs of the UNMODIFIED reference WanModel (tests/golden/make_model_golden.py), fp32 route.  CPU."""
import os

import pytest
import torch

from oracle import ref_loader
from oracle import wan_attention_oracle as orc
from tests.golden import make_model_golden as mg

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wan_model_golden.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, map_location="cpu", weights_only=False)


def _state():
    """seeded_state only needs the key -> shape map; take it from the drop-in model (same names as the reference,
    tests/test_reference_compat.py::test_parameter_names_match_reference_modules)."""
    from univid_b200.wan.modules import model as mine
    return mg.seeded_state(mine.WanModel(**mg.KW))


def test_golden_inputs_have_not_drifted(gold):
    got = mg.checksums(mg.model_case())
    for k, v in gold["checksums"].items():
        assert abs(got[k] - v) <= 1e-9 * max(1.0, abs(v)), k
    assert gold["kw"] == mg.KW and gold["seq_len"] == mg.SEQ_LEN


@pytest.mark.parametrize("form", ["scalar", "token"])
def test_oracle_dit_forward_matches_reference_golden(gold, form):
    case, prm, kw = mg.model_case(), _state(), mg.KW
    lat, t, ctx = mg.inputs_for(case, form)
    got = orc.dit_forward(lat, t, ctx, prm, case["seq_len"], kw["num_heads"], kw["num_layers"], kw["dim"], kw["freq_dim"],
                          kw["text_len"], kw["out_dim"], eps=kw["eps"], bf16=False)
    want = gold[f"{form}_fp32"]
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a.shape == b.shape
        assert (a - b).abs().max().item() <= 2e-5 * max(1.0, b.abs().max().item())   # fp32 op-order differences only


@pytest.mark.skipif(not ref_loader.available(), reason="reference sources not available")
@pytest.mark.parametrize("form", ["scalar", "token"])
def test_live_reference_reproduces_the_golden(gold, form):
    import warnings
    att, model = ref_loader.load_modules()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = mg.run_reference_fp32(model, att, mg.model_case(), form)
    for a, b in zip(out, gold[f"{form}_fp32"]):
        assert torch.equal(a, b)
