"""ProteInfer.from_pretrained (protein_encoders.py:125-153) goes through the reference's own positional
TF -> torch transfer (protnote/utils/proteinfer.py:7-41), which zips the TensorFlow variables onto `state_dict()` ORDER.
A synthetic TF-variable pickle is loaded into the reference class and into protnote_b200's; the two state_dicts must be
identical entry by entry (names, order, shapes, values).  CPU only; needs the reference tree for the transfer utility."""
import pickle

import numpy as np
import pytest
import torch

from oracle.ref_import import import_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present on this machine")

CFG = dict(num_labels=7, input_channels=20, output_channels=24, kernel_size=9, activation=torch.nn.ReLU,
           dilation_base=3, num_resnet_blocks=3, bottleneck_factor=0.5)


def synthetic_tf_variables(cfg, seed=0):
    """TF checkpoint variables in graph order, as `ProteInfer`'s TF original lays them out: Conv1D kernels are
    (k, Cin, Cout), Dense kernels (in, out), BatchNorm gamma/beta/moving_mean/moving_variance; one global step."""
    rng = np.random.default_rng(seed)
    C, Cb, k = cfg["output_channels"], int(cfg["output_channels"] * cfg["bottleneck_factor"]), cfg["kernel_size"]
    f = lambda *s: rng.standard_normal(s).astype(np.float32)  # noqa: E731
    tf = {"inferrer/conv1d/kernel:0": f(k, cfg["input_channels"], C), "inferrer/conv1d/bias:0": f(C)}
    n = 0
    for i in range(cfg["num_resnet_blocks"]):
        for width, (kk, cin, cout) in ((C, (k, C, Cb)), (Cb, (1, Cb, C))):
            bn = f"inferrer/residual_block_{i}/batch_normalization_{n}"
            tf[bn + "/gamma:0"], tf[bn + "/beta:0"] = f(width), f(width)
            tf[bn + "/moving_mean:0"], tf[bn + "/moving_variance:0"] = f(width), np.abs(f(width)) + 0.5
            n += 1
            tf[f"inferrer/residual_block_{i}/conv1d_{n}/kernel:0"] = f(kk, cin, cout)
            tf[f"inferrer/residual_block_{i}/conv1d_{n}/bias:0"] = f(cout)
    tf["inferrer/logits/kernel:0"], tf["inferrer/logits/bias:0"] = f(C, cfg["num_labels"]), f(cfg["num_labels"])
    tf["inferrer/global_step:0"] = np.int64(12345)
    return tf


def test_from_pretrained_matches_reference(tmp_path):
    _, RefProteInfer, _ = import_reference()
    from protnote_b200.protein_encoders import ProteInfer
    path = tmp_path / "tf_weights.pkl"
    with open(path, "wb") as fh:
        pickle.dump(synthetic_tf_variables(CFG), fh)
    ref = RefProteInfer.from_pretrained(weights_path=str(path), **CFG)
    ours = ProteInfer.from_pretrained(weights_path=str(path), **CFG)
    sd_ref, sd_ours = ref.state_dict(), ours.state_dict()
    assert list(sd_ref.keys()) == list(sd_ours.keys())          # the transfer is positional: order IS the contract
    for name in sd_ref:
        assert sd_ref[name].shape == sd_ours[name].shape, name
        assert sd_ref[name].dtype == sd_ours[name].dtype, name
        assert torch.equal(sd_ref[name], sd_ours[name]), name
    # the values really landed where the reference puts them (not just "both are untouched random init")
    tf = synthetic_tf_variables(CFG)
    assert torch.equal(sd_ours["conv1.weight"], torch.from_numpy(tf["inferrer/conv1d/kernel:0"].transpose(2, 1, 0)))
    assert torch.equal(sd_ours["resnet_blocks.2.masked_conv2.bias"],
                       torch.from_numpy(tf["inferrer/residual_block_2/conv1d_6/bias:0"]))
    assert int(sd_ours["resnet_blocks.0.bn_activation_1.0.num_batches_tracked"]) == 12345
    assert torch.equal(sd_ours["output_layer.weight"], torch.from_numpy(tf["inferrer/logits/kernel:0"].T))


def test_transfer_rejects_a_reordered_encoder(tmp_path):
    """If the parameter registration order drifted, the positional transfer would hit a shape mismatch (bottleneck vs
    full width); the reference asserts on it - and so does the drop-in, because it calls the same utility."""
    from protnote_b200.protein_encoders import ProteInfer
    import_reference()
    tf = synthetic_tf_variables(CFG)
    keys = list(tf)
    i, j = keys.index("inferrer/residual_block_0/conv1d_1/kernel:0"), keys.index("inferrer/residual_block_0/conv1d_2/kernel:0")
    keys[i], keys[j] = keys[j], keys[i]
    path = tmp_path / "tf_weights_bad.pkl"
    with open(path, "wb") as fh:
        pickle.dump({k: tf[k] for k in keys}, fh)
    with pytest.raises(AssertionError):
        ProteInfer.from_pretrained(weights_path=str(path), **CFG)
