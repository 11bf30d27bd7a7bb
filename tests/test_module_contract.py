"""Constructor / state_dict / error-behaviour parity of the drop-in modules (CPU-only checks)."""
import pytest
import torch

import dualmessagepassing_b200 as dmp
from tests import _golden

SEEDS = {"scm_rev_mlp2_lrelu": 7000, "scm_rev_mlp0_relu": 7001, "scm_norev_mlp2_bn_tanh": 7002,
         "scm_rev_din_ne_h_nobias": 7003, "scm_rev_h64_medium": 7004, "scm_rev_mlp1_h50": 7005}
UNC_SEEDS = {"unc_norm_bn_tanh": 7200, "unc_norm_bn_last": 7201, "unc_isrev_nobn": 7202, "unc_h50_medium": 7203}


@pytest.mark.parametrize("name", sorted(SEEDS))
def test_dmplayer_init_is_seed_identical_to_reference(name):
    case = _golden.load(name)
    din, h, mlp, bn, bias = [int(x) for x in case["meta"]]
    torch.manual_seed(SEEDS[name])
    layer = dmp.DMPLayer(din, h, bias=bool(bias), num_mlp_layers=mlp, batch_norm=bool(bn), act_func=case["act"])
    sd = layer.state_dict()
    assert set(sd) == set(case["params"])
    for k, v in sd.items():
        assert v.shape == case["params"][k].shape, k
        assert torch.equal(v, case["params"][k]), k


@pytest.mark.parametrize("name", sorted(UNC_SEEDS))
def test_dualgraphconv_init_is_seed_identical_to_reference(name):
    case = _golden.load(name)
    din, h, _, bn, _ = [int(x) for x in case["meta"]]
    torch.manual_seed(UNC_SEEDS[name])
    act = torch.nn.Tanh() if case["act"] == "tanh" else None
    layer = dmp.DualGraphConv(din, h, batch_norm=bool(bn), activation=act)
    sd = layer.state_dict()
    assert set(sd) == set(case["params"])  # includes the unused nfc/efc of model.py:137-138
    for k, v in sd.items():
        assert torch.equal(v, case["params"][k]), k


@pytest.mark.parametrize("name,seed", [("lrp_h16_mlp2", 7300), ("lrp_h12_mlp0_bn", 7301)])
def test_dmplrp_init_is_seed_identical_to_reference(name, seed):
    case = _golden.load(name)
    h, L, D, mlp, bn = [int(x) for x in case["meta"]]
    torch.manual_seed(seed)
    layer = dmp.DMPLRPPoolLayer(h, h, lrp_seq_len=L, num_mlp_layers=mlp, batch_norm=bool(bn), act_func=case["act"])
    sd = layer.state_dict()
    assert set(sd) == set(case["params"])      # lrp_weight [in, hid, L*L], lrp_bias (dmplrp.py:45-53)
    for k, v in sd.items():
        assert torch.equal(v, case["params"][k]), k


def test_state_dict_roundtrip_and_registered_none_bias():
    layer = dmp.DMPLayer(8, 12, bias=False, num_mlp_layers=2, batch_norm=True, act_func="relu")
    assert layer.nbias is None and layer.ebias is None
    assert "nmlp.1.running_mean" in layer.state_dict() and "nmlp.3.weight" in layer.state_dict()
    other = dmp.DMPLayer(8, 12, bias=False, num_mlp_layers=2, batch_norm=True, act_func="relu")
    other.load_state_dict(layer.state_dict())
    assert layer.get_output_dim() == 12 and "in=8, out=12" in repr(layer)


def test_activation_modules_are_shared_singletons():
    a = dmp.DMPLayer(4, 4, act_func="leaky_relu")
    b = dmp.DMPLayer(4, 4, act_func="leaky_relu")
    assert a.act is b.act and a.nmlp[2] is b.emlp[2]  # utils/act.py:457-489 hands out one instance per name
    with pytest.raises(NotImplementedError):
        dmp.DMPLayer(4, 4, act_func="sparsemax")


def test_cpu_tensors_fail_loudly_no_fallback():
    g = dmp.DMPGraph([0, 1], [1, 0], 2)
    layer = dmp.DMPLayer(4, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        layer(g, torch.randn(2, 4), torch.randn(2, 4))


def test_missing_library_fails_loudly(monkeypatch):
    from dualmessagepassing_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libdmp_b200.so")
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.load()
