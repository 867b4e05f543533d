"""The C-ABI library loads on a machine without a GPU and exports every symbol include/protnote_b200.h declares;
host-side argument checking works without a device (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

from protnote_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()          # no-op when the in-tree .so is newer than its sources
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    header = open(os.path.join(ROOT, "include", "protnote_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(pn_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    for name in declared:
        assert getattr(lib, name) is not None


def test_sizes_and_argument_errors_without_a_device(lib):
    enc = _lib.EncoderCfg(20, 1100, 550, 9, 3, 5, 1e-3)
    assert lib.pn_encoder_packed_bytes(C.byref(enc)) > 2 * (1100 * 9 * 1152 * 550 // 550) * 2
    assert lib.pn_encoder_workspace_bytes(C.byref(enc), 2, 1024) > 2 * 1024 * 1100 * 4
    bad = _lib.EncoderCfg(20, 1100, 550, 8, 3, 5, 1e-3)           # even kernel size
    assert lib.pn_encoder_packed_bytes(C.byref(bad)) == 0
    assert b"odd" in lib.pn_last_error()
    sc = _lib.ScorerCfg(1100, 1024, 1024, 3072, 4, 3072, 3, 1, 0, 1, 1e-5)
    assert lib.pn_scorer_num_params(C.byref(sc)) == 49
    assert lib.pn_scorer_packed_bytes(C.byref(sc)) > 3 * 3072 * 3072 * 4
    assert lib.pn_scorer_min_workspace_bytes(C.byref(sc)) > 0
    sc_one = _lib.ScorerCfg(1100, 1024, 1024, 3072, 4, 3072, 1, 1, 0, 1, 1e-5)       # OUTPUT_MLP_NUM_LAYERS 1
    assert lib.pn_scorer_num_params(C.byref(sc_one)) == 2 * 16 + 5 + 2
    sc_bad = _lib.ScorerCfg(1100, 1024, 1024, 3072, 4, 3072, 0, 1, 0, 1, 1e-5)
    assert lib.pn_scorer_num_params(C.byref(sc_bad)) == -1 and b"out_layers" in lib.pn_last_error()
    assert lib.pn_set_option(b"bk", 48) != 0 and lib.pn_set_option(b"no_such_option", 1) != 0
    assert lib.pn_set_option(b"bk", 0) == 0


def test_module_interface_matches_reference_contract():
    """Constructor keywords, state_dict key names and error behaviour of the host-side mirror."""
    import torch
    from protnote_b200.ProtNote import ProtNote
    from protnote_b200.protein_encoders import ProteInfer
    enc = ProteInfer(num_labels=7, input_channels=20, output_channels=24, kernel_size=9, activation=torch.nn.ReLU,
                     dilation_base=3, num_resnet_blocks=2, bottleneck_factor=0.5)
    model = ProtNote(protein_embedding_dim=24, label_embedding_dim=16, latent_dim=8, sequence_encoder=enc,
                     label_encoder=None, output_mlp_hidden_dim_scale_factor=3, output_mlp_num_layers=3,
                     outout_mlp_add_batchnorm=True, projection_head_num_layers=4,
                     projection_head_hidden_dim_scale_factor=3, feature_fusion="concatenation")
    keys = list(model.state_dict())
    for want in ("sequence_encoder.conv1.weight", "sequence_encoder.resnet_blocks.1.bn_activation_2.0.running_var",
                 "sequence_encoder.resnet_blocks.0.masked_conv1.bias", "sequence_encoder.output_layer.weight",
                 "W_p.0.weight", "W_p.1.running_mean", "W_p.12.weight", "W_l.8.weight", "output_layer.0.weight",
                 "output_layer.9.num_batches_tracked", "output_layer.11.bias"):
        assert want in keys, want
    # parameter registration order of the encoder = the reference's (transfer_tf_weights_to_torch zips by position)
    enc_keys = [k for k in enc.state_dict()]
    assert enc_keys[:2] == ["conv1.weight", "conv1.bias"]
    assert enc_keys[2].startswith("resnet_blocks.0.bn_activation_1.0.") and enc_keys[-2:] == ["output_layer.weight", "output_layer.bias"]
    dropped = ProtNote(protein_embedding_dim=24, label_embedding_dim=16, latent_dim=8, sequence_embedding_dropout=0.1,
                       projection_head_num_layers=2, projection_head_hidden_dim_scale_factor=2,
                       output_mlp_hidden_dim_scale_factor=2, output_mlp_num_layers=2)
    assert "W_p.1.0.weight" in dropped.state_dict() and "W_l.0.weight" in dropped.state_dict()
    model.eval()
    x = torch.zeros(2, 20, 12)
    with pytest.raises(ValueError):
        model(sequence_onehots=x, sequence_lengths=torch.tensor([12, 7]))           # no labels
    with pytest.raises(_lib.ProtnoteB200Error):                                         # no CPU path
        model(sequence_onehots=x, sequence_lengths=torch.tensor([12, 7]), label_embeddings=torch.zeros(4, 16))
    model.train()
    with pytest.raises(_lib.ProtnoteB200Error):
        model(sequence_onehots=x, sequence_lengths=torch.tensor([12, 7]), label_embeddings=torch.zeros(4, 16))


def test_header_is_plain_c(tmp_path):
    """include/protnote_b200.h is the boundary a non-Python host binds: it must compile as C99 on its own, and a C
    translation unit that calls the entry points must link against the shared library."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "host.c"
    src.write_text(
        '#include "protnote_b200.h"\n'
        '#include <stdio.h>\n'
        'int main(void) {\n'
        '  pn_scorer_cfg cfg = {1100, 1024, 1024, 3072, 4, 3072, 3, 1, PN_FUSION_CONCAT, 1, 1e-5f};\n'
        '  pn_encoder_cfg enc = {20, 1100, 550, 9, 3, 5, 1e-3f};\n'
        '  printf("%d %zu %zu %d\\n", pn_version(), pn_scorer_packed_bytes(&cfg), pn_encoder_packed_bytes(&enc),\n'
        '         pn_scorer_num_params(&cfg));\n'
        '  return pn_score_pairs(&cfg, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, PN_STRICT, 0) == 0;   /* empty input -> error status */\n'
        '}\n')
    exe = tmp_path / "host"
    lib_dir = os.path.join(root, "protnote_b200", "lib")
    res = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                          "-L", lib_dir, "-lprotnote_b200", "-Wl,-rpath," + lib_dir], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0, run.stderr           # sizes are host arithmetic; the empty call returns an error status
    version, scorer_bytes, encoder_bytes, nparams = run.stdout.split()
    assert int(version) >= 1 and int(scorer_bytes) > 100e6 and int(encoder_bytes) > 50e6 and int(nparams) == 49


def test_label_fingerprint_digest_is_order_sensitive():
    """The on-disk label-projection cache is keyed by a digest of the label embeddings and weights: permuting rows (a
    re-sorted vocabulary) or columns must change it, an identical copy must not."""
    import torch
    from protnote_b200.ProtNote import ProtNote
    g = torch.Generator().manual_seed(0)
    t = torch.randn(37, 16, generator=g)
    base = ProtNote._digest(t)
    assert ProtNote._digest(t.clone()) == base
    perm = torch.randperm(37, generator=g)
    assert ProtNote._digest(t[perm])[:2] == pytest.approx(base[:2], rel=1e-12)      # plain sums cannot see a permutation
    assert abs(ProtNote._digest(t[perm])[2] - base[2]) > 1e-6 * abs(base[2])        # the position-weighted sum does
    assert abs(ProtNote._digest(t[:, torch.randperm(16, generator=g)])[2] - base[2]) > 1e-6 * abs(base[2])


def test_token_ids_are_range_checked_before_narrowing():
    """CPU token tensors are narrowed to uint8 for the PCIe copy: ids outside [0, 255] (a -1 padding id, 256+) must raise
    instead of wrapping into a valid residue - before any device work."""
    import torch
    from protnote_b200.protein_encoders import ProteInfer
    enc = ProteInfer(num_labels=4, input_channels=20, output_channels=16, kernel_size=9, activation=torch.nn.ReLU,
                     dilation_base=3, num_resnet_blocks=1, bottleneck_factor=0.5).eval()
    lengths = torch.tensor([5, 5])
    for bad in (torch.tensor([[0, 1, -1, 3, 4], [0, 1, 2, 3, 4]]), torch.tensor([[0, 1, 256, 3, 4], [0, 1, 2, 3, 4]]),
                torch.rand(2, 5)):
        with pytest.raises(ValueError):
            enc.get_embeddings_from_tokens(bad, lengths)
