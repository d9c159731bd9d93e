"""-m "not gpu": the flat parameter layout carries exactly the reference's state_dict names and shapes."""
import torch

from tests.helpers import Golden


def test_state_dict_roundtrip_matches_reference_names():
    from spvipes_b200.engine import StepEngine
    gd = Golden("label_tiny")
    eng = StepEngine((gd.G0, gd.G1), gd.H, gd.S, gd.P, gd.dropout, gd.mode, device="cpu")
    eng.load_state_dict(gd.sd)
    sd = eng.state_dict()
    ref_names = [k for k in gd.sd if not k.endswith("num_batches_tracked")]
    assert sorted(sd.keys()) == sorted(ref_names)
    for k in ref_names:
        assert tuple(sd[k].shape) == tuple(gd.sd[k].shape), k
        assert torch.equal(sd[k], gd.sd[k].float()), k
    # fused blocks are contiguous concatenations of the per-encoder tensors
    H = gd.H
    W1 = eng.P(0, "W1")
    assert torch.equal(W1[:H], gd.sd["encoder_0_private.fc1.weight"]) and torch.equal(W1[H:], gd.sd["encoder_0_shared.fc1.weight"])
    assert set(eng.grad_dict().keys()) == set(gd.grads.keys())
