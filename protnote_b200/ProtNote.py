"""B200-native mirror of the reference's `protnote/models/ProtNote.py`.

`ProtNote` keeps the reference constructor (ProtNote.py:10-36, incl. the `outout_mlp_add_batchnorm` spelling),
`forward()` signature and return value (:168-177,324-334), submodule / state_dict names (`W_p.{0,4,8,12}`,
`W_l.*`, `output_layer.{0,1,4,5,8,9,11}`, `sequence_encoder.*`, `label_encoder.*`) and error behaviour
(ValueError on incompatible arguments, :217,262-264,305), so it drops into `bin/main.py:407-452` and
`ProtNoteTrainer.evaluation_step` (ProtNoteTrainer.py:247-292) unchanged.

What differs is where the arithmetic happens: eval-mode forward = sm_100a kernels behind the C ABI
  sequence_encoder.get_embeddings -> pn_encoder_forward
  W_p (+ protein half of output layer 1) -> pn_project_sequences
  W_l (+ label half of output layer 1)   -> pn_project_labels   (cached across batches: constant in eval mode)
  joint features + output MLP + ensembling -> pn_score_pairs    (the [B*L, 2d] joint tensor never exists)
There is no PyTorch/CPU implementation of that path here; unsupported modes raise.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import native
from ._lib import ProtnoteB200Error
from .protein_encoders import _versions


def _projection_mlp(in_dim: int, hidden_dims, dropout: float) -> nn.Sequential:
    """Same module sequence (hence the same state_dict indices) as torchvision.ops.MLP(in, hidden, bias=False,
    norm_layer=BatchNorm1d, dropout=dropout) used at ProtNote.py:63-81: [Linear, BN, ReLU, Dropout]*, Linear, Dropout."""
    layers = []
    d = in_dim
    for h in hidden_dims[:-1]:
        layers += [nn.Linear(d, h, bias=False), nn.BatchNorm1d(h), nn.ReLU(), nn.Dropout(dropout)]
        d = h
    layers += [nn.Linear(d, hidden_dims[-1], bias=False), nn.Dropout(dropout)]
    return nn.Sequential(*layers)


def get_mlp(input_dim, hidden_dim, num_layers, input_dropout=0.0, dropout=0.0, batch_norm=False,
            output_neuron_bias=None):
    """Parameter container with the reference's layer order (ProtNote.py:337-378)."""
    layers = []
    if input_dropout > 0:
        layers.append(nn.Dropout(input_dropout))
    for idx in range(num_layers):
        layers.append(nn.Linear(input_dim if idx == 0 else hidden_dim, hidden_dim, bias=not batch_norm))
        if batch_norm:
            layers.append(nn.BatchNorm1d(hidden_dim))
        layers.append(nn.ReLU())
        if idx < num_layers - 1:
            layers.append(nn.Dropout(dropout))
    output_neuron = nn.Linear(hidden_dim, 1)
    if output_neuron_bias is not None:
        output_neuron.bias.data.fill_(output_neuron_bias)
    layers.append(output_neuron)
    return nn.Sequential(*layers)


class ProtNote(nn.Module):
    def __init__(
        self,
        protein_embedding_dim=1100,
        label_embedding_dim=1024,
        label_embedding_pooling_method="mean",
        inference_descriptions_per_label=1,
        latent_dim=1024,
        label_encoder=None,
        sequence_encoder=None,
        label_encoder_num_trainable_layers=False,
        train_sequence_encoder=False,
        output_mlp_hidden_dim_scale_factor=1024,
        output_mlp_num_layers=2,
        output_neuron_bias=None,
        outout_mlp_add_batchnorm=True,
        residual_connection=False,
        dropout=0.0,
        sequence_embedding_dropout=0.0,
        label_embedding_dropout=0.0,
        label_embedding_noising_alpha=0.0,
        projection_head_num_layers=1,
        projection_head_hidden_dim_scale_factor=1,
        label_batch_size_limit=float("inf"),
        sequence_batch_size_limit=float("inf"),
        feature_fusion="concatenation",
        temperature=0.07,
        precision="strict",
    ):
        super().__init__()
        self.label_encoder_num_trainable_layers, self.train_sequence_encoder = (
            label_encoder_num_trainable_layers, train_sequence_encoder)
        self.label_encoder, self.sequence_encoder = label_encoder, sequence_encoder
        self.inference_descriptions_per_label = inference_descriptions_per_label
        self.label_batch_size_limit, self.sequence_batch_size_limit = label_batch_size_limit, sequence_batch_size_limit
        self.feature_fusion = feature_fusion
        self.temperature = temperature
        self.label_embedding_pooling_method = label_embedding_pooling_method
        self.latent_dim = latent_dim
        self.label_embedding_noising_alpha = label_embedding_noising_alpha
        self.residual_connection = residual_connection
        self.precision = precision
        # The reference's last Linear runs under torch.autocast in ProtNoteTrainer.evaluation_step (:287-290), so ITS
        # logits come back fp16 there; this module computes fp32-grade logits regardless of autocast and returns fp32.
        # Set True to cast the returned logits to the active autocast dtype (dtype drop-in for callers that depend on it).
        self.match_autocast_dtype = False

        hidden = [latent_dim * projection_head_hidden_dim_scale_factor] * (projection_head_num_layers - 1) + [latent_dim]
        self.W_p = _projection_mlp(protein_embedding_dim, hidden, dropout)
        self.W_l = _projection_mlp(label_embedding_dim, hidden, dropout)
        # the reference wraps the heads when these are > 0, which renames the keys to W_p.1.* / W_l.1.* (:83-86)
        if sequence_embedding_dropout > 0:
            self.W_p = nn.Sequential(nn.Dropout(sequence_embedding_dropout), self.W_p)
        if label_embedding_dropout > 0:
            self.W_l = nn.Sequential(nn.Dropout(label_embedding_dropout), self.W_l)
        if self.label_embedding_pooling_method == "all":
            self.raw_attn_scorer = nn.Linear(label_embedding_dim, 1, bias=True)
        if self.feature_fusion.startswith("concatenation"):
            self.output_layer = get_mlp(
                input_dim=self._get_concatenated_features_dim(),
                hidden_dim=int(round(output_mlp_hidden_dim_scale_factor * latent_dim)),
                num_layers=output_mlp_num_layers,
                output_neuron_bias=output_neuron_bias,
                batch_norm=outout_mlp_add_batchnorm,
                dropout=dropout,
            )
        self._cfg = dict(protein_dim=protein_embedding_dim, label_dim=label_embedding_dim, latent_dim=latent_dim,
                         proj_hidden=latent_dim * projection_head_hidden_dim_scale_factor,
                         proj_layers=projection_head_num_layers,
                         out_hidden=int(round(output_mlp_hidden_dim_scale_factor * latent_dim)),
                         out_layers=output_mlp_num_layers, out_batchnorm=bool(outout_mlp_add_batchnorm))
        self._packed = None
        self._packed_key = None
        self._label_cache = None

    def _get_concatenated_features_dim(self):
        dim = {"concatenation_diff": self.latent_dim * 3, "concatenation_prod": self.latent_dim * 3,
               "concatenation": self.latent_dim * 2}
        return dim[self.feature_fusion]

    # ------------------------------------------------------------------ packed-weight cache
    @staticmethod
    def _head_sources(head: nn.Module):
        if isinstance(head[0], nn.Dropout) and isinstance(head[1], nn.Sequential):
            head = head[1]
        srcs, mods = [], list(head)
        for i, m in enumerate(mods):
            if isinstance(m, nn.Linear):
                srcs.append(m.weight)
                if i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm1d):
                    bn = mods[i + 1]
                    srcs += [bn.weight, bn.bias, bn.running_mean, bn.running_var]
        return srcs

    def _pack_sources(self):
        srcs = self._head_sources(self.W_p) + self._head_sources(self.W_l)
        if self.feature_fusion == "similarity":
            return srcs
        mods = list(self.output_layer)
        for i, m in enumerate(mods[:-1]):
            if isinstance(m, nn.Linear):
                srcs.append(m.weight)
                if isinstance(mods[i + 1], nn.BatchNorm1d):
                    bn = mods[i + 1]
                    srcs += [bn.weight, bn.bias, bn.running_mean, bn.running_var]
                else:
                    srcs.append(m.bias)
        srcs += [mods[-1].weight, mods[-1].bias]
        return srcs

    def _ensure_packed(self):
        srcs = self._pack_sources()
        key = _versions(srcs) + (self.feature_fusion, self.inference_descriptions_per_label, native.options_epoch())
        if self._packed is None or key != self._packed_key:
            own = [self.W_p, self.W_l] + ([self.output_layer] if hasattr(self, "output_layer") else [])
            bn_eps = next((m.eps for part in own for m in part.modules() if isinstance(m, nn.BatchNorm1d)), 1e-5)
            sc = native.PackedScorer(fusion=self.feature_fusion,
                                     descriptions_per_label=self.inference_descriptions_per_label, bn_eps=bn_eps,
                                     **self._cfg)
            sc.pack(srcs)
            self._packed, self._packed_key, self._label_cache = sc, key, None
        return self._packed

    def _projected_labels(self, scorer, L_f, mode, want_embedding):
        key = (L_f.data_ptr(), L_f._version, tuple(L_f.shape), self._packed_key, mode, want_embedding)
        if self._label_cache is not None and self._label_cache[0] == key:
            return self._label_cache[1]
        out = scorer.project_labels(L_f, mode, want_embedding=want_embedding)
        # keep a reference to L_f so its storage (the cache key) cannot be recycled for other data
        self._label_cache = (key, out, L_f)
        return out

    # ------------------------------------------------------------------ label-projection cache on disk
    @staticmethod
    def _digest(t):
        """Order-sensitive device-side checksum of a tensor: plain sum, sum of magnitudes and a sum weighted by a fixed
        pseudo-random function of the element POSITION (a permutation of rows, columns or entries changes it)."""
        v = t.detach().double().flatten()
        idx = torch.arange(v.numel(), dtype=torch.int64, device=v.device)
        w = ((idx * 2654435761) & 0xFFFFFFFF).double() / 4294967296.0 + 0.5
        return [float(v.sum()), float(v.abs().sum()), float((v * w).sum())]

    def _label_fingerprint(self, L_f):
        """What the projected label halves depend on: W_l, layer 1 of the output MLP (+ its BatchNorm, incl. eps), the
        label embeddings themselves IN ORDER (rows are labels: a re-sorted vocabulary must not match), the number of
        descriptions per label and the arithmetic mode."""
        parts = list(self._head_sources(self.W_l))
        eps = [m.eps for m in self.W_l.modules() if isinstance(m, nn.BatchNorm1d)]
        if self.feature_fusion != "similarity":
            mods = list(self.output_layer)
            parts.append(mods[0].weight)
            if isinstance(mods[1], nn.BatchNorm1d):
                parts += [mods[1].weight, mods[1].bias, mods[1].running_mean, mods[1].running_var]
                eps.append(mods[1].eps)
            elif mods[0].bias is not None:
                parts.append(mods[0].bias)
        sums = [x for t in parts for x in self._digest(t)]
        return {"weights": sums, "labels": self._digest(L_f), "label_shape": list(L_f.shape), "precision": self.precision,
                "fusion": self.feature_fusion, "bn_eps": [float(e) for e in eps],
                "descriptions_per_label": int(self.inference_descriptions_per_label)}

    def save_label_projection(self, path, label_embeddings):
        """Runs W_l and the label half of output layer 1 once for `label_embeddings` [L, label_dim] and stores the result
        (c [L, H] and, for the fusions that need it, L_e [L, latent]) with a fingerprint of everything it depends on.  The
        cached-embedding file of the reference (bin/generate_label_embeddings.py:149-164) only holds the raw text
        embeddings, so W_l is recomputed by every process and every batch (ProtNote.py:271)."""
        if self.training:
            raise ProtnoteB200Error("label projections are cached for evaluation (model.eval())")
        dev = next(self.W_p.parameters()).device
        L_f = label_embeddings.to(dev)
        scorer = self._ensure_packed()
        need_emb = self.feature_fusion in ("concatenation_prod", "similarity")
        L_e, c = scorer.project_labels(L_f, native.MODES[self.precision], want_embedding=need_emb)
        torch.save({"fingerprint": self._label_fingerprint(L_f), "c": None if c is None else c.cpu(),
                    "L_e": None if L_e is None else L_e.cpu()}, path)

    def load_label_projection(self, path, label_embeddings):
        """Installs a stored projection for `label_embeddings` (same tensor object must then be passed to forward());
        raises if the weights, the embeddings or the mode differ from the ones it was computed with."""
        dev = next(self.W_p.parameters()).device
        L_f = label_embeddings.to(dev)
        blob = torch.load(path)
        want = self._label_fingerprint(L_f)
        got = blob["fingerprint"]
        same = (got["label_shape"] == want["label_shape"] and got["precision"] == want["precision"]
                and got["fusion"] == want["fusion"] and got.get("bn_eps") == want["bn_eps"]
                and got.get("descriptions_per_label") == want["descriptions_per_label"]
                and len(got["weights"]) == len(want["weights"]) and len(got["labels"]) == len(want["labels"])
                and all(abs(a - b) <= 1e-9 * max(1.0, abs(b)) for a, b in zip(got["weights"] + got["labels"],
                                                                                want["weights"] + want["labels"])))
        if not same:
            raise ProtnoteB200Error("stored label projection does not match this model / these label embeddings")
        self._ensure_packed()
        mode = native.MODES[self.precision]
        need_emb = self.feature_fusion in ("concatenation_prod", "similarity")
        out = (None if blob["L_e"] is None else blob["L_e"].to(dev), None if blob["c"] is None else blob["c"].to(dev))
        for want_embedding in {need_emb, True} if out[0] is not None else {need_emb}:
            key = (L_f.data_ptr(), L_f._version, tuple(L_f.shape), self._packed_key, mode, want_embedding)
            self._label_cache = (key, out, L_f)
        return L_f

    # ------------------------------------------------------------------ training mode
    def _forward_train(self, sequence_onehots, sequence_embeddings, sequence_lengths, label_embeddings,
                       label_token_counts, save_embeddings):
        """ProtNote.forward with self.training (ProtNote.py:168-334): BatchNorm batch statistics in W_p / W_l / output MLP,
        label-embedding noise, logits [B, L] differentiable w.r.t. W_p, W_l and output_layer.  The arithmetic is the
        `pn_t_*` primitives sequenced by protnote_b200/train.py (custom autograd.Function)."""
        from . import train as pn_train
        if save_embeddings:
            raise ProtnoteB200Error("save_embeddings=True is an evaluation diagnostic; it is not available in training mode")
        if self.train_sequence_encoder or self.label_encoder_num_trainable_layers:
            raise ProtnoteB200Error("the sm_100a training path trains W_p, W_l and output_layer; the sequence encoder and the "
                                    "label text encoder are frozen (TRAIN_SEQUENCE_ENCODER False, "
                                    "LABEL_ENCODER_NUM_TRAINABLE_LAYERS 0: base_config.yaml:71-72)")
        if label_embeddings is None:
            raise ValueError("Incompatible label parameters passed to forward method.")
        if self.label_embedding_pooling_method == "all":
            raise ProtnoteB200Error("LABEL_EMBEDDING_POOLING_METHOD 'all' is not on the cached-embedding path")
        if not (self.feature_fusion.startswith("concatenation") or self.feature_fusion == "similarity"):
            raise ValueError("feature fusion method not implemented")
        dev = next(self.W_p.parameters()).device
        L_f = label_embeddings.to(dev, non_blocking=True)
        # label-embedding noise (ProtNote.py:219-240): RNG-stream dependent, kept as the reference's own torch ops
        if label_token_counts is not None and self.label_embedding_noising_alpha > 0:
            scalars = self.label_embedding_noising_alpha / math.sqrt(L_f.shape[1])
            L_f = L_f + (2 * torch.rand_like(L_f) - 1) * scalars
        if sequence_embeddings is not None:
            P_f = sequence_embeddings.to(dev, non_blocking=True)
        elif sequence_onehots is not None and sequence_lengths is not None:
            if self.sequence_encoder is None:
                raise ValueError("Incompatible sequence parameters passed to forward method.")
            with torch.no_grad():
                P_f = self.sequence_encoder.get_embeddings(sequence_onehots.to(dev, non_blocking=True),
                                                           sequence_lengths.to(dev, non_blocking=True))
        else:
            raise ValueError("Incompatible sequence parameters passed to forward method.")
        logits = pn_train.train_logits(self, P_f, L_f)
        # BatchNorm running statistics were updated by the kernels through raw pointers (and the optimizer is about to
        # change the weights): the eval-mode pack and the cached label projection are stale from here on
        self._packed_key = None
        self._label_cache = None
        return logits, {"output_layer_embeddings": [], "joint_embeddings": []}

    # ------------------------------------------------------------------ reference interface
    def forward(self, sequence_onehots=None, sequence_embeddings=None, sequence_lengths=None, tokenized_labels=None,
                label_embeddings=None, label_token_counts=None, save_embeddings=False, sequence_tokens=None):
        """Reference signature (ProtNote.py:168-177) plus one optional extension: `sequence_tokens` [B, T] integer residue ids
        may replace `sequence_onehots` (same result bit for bit, 80x less host-to-device traffic)."""
        if sequence_tokens is not None and sequence_onehots is None and sequence_embeddings is None:
            if self.sequence_encoder is None or sequence_lengths is None:
                raise ValueError("Incompatible sequence parameters passed to forward method.")
            with torch.no_grad():
                sequence_embeddings = self.sequence_encoder.get_embeddings_from_tokens(sequence_tokens, sequence_lengths)
        if self.training:
            return self._forward_train(sequence_onehots, sequence_embeddings, sequence_lengths, label_embeddings,
                                       label_token_counts, save_embeddings)
        # ---- label embeddings (ProtNote.py:192-217): cached embeddings only
        if label_embeddings is not None:
            L_f = label_embeddings
        else:
            raise ValueError("Incompatible label parameters passed to forward method.")
        if self.label_embedding_pooling_method == "all":
            raise ProtnoteB200Error("LABEL_EMBEDDING_POOLING_METHOD 'all' (token-level attention pooling, "
                                    "ProtNote.py:154-166) is not on the cached-embedding path")
        if not (self.feature_fusion.startswith("concatenation") or self.feature_fusion == "similarity"):
            raise ValueError("feature fusion method not implemented")
        dev = next(self.W_p.parameters()).device
        mode = native.MODES[self.precision]
        scorer = self._ensure_packed()
        # ---- sequence embeddings (ProtNote.py:243-264)
        if sequence_embeddings is not None:
            P_f = sequence_embeddings.to(dev, non_blocking=True)
        elif sequence_onehots is not None and sequence_lengths is not None:
            if self.sequence_encoder is None:
                raise ValueError("Incompatible sequence parameters passed to forward method.")
            P_f = self.sequence_encoder.get_embeddings(sequence_onehots.to(dev, non_blocking=True),
                                                       sequence_lengths.to(dev, non_blocking=True))
        else:
            raise ValueError("Incompatible sequence parameters passed to forward method.")
        L_f = L_f.to(dev, non_blocking=True)
        need_emb = self.feature_fusion in ("concatenation_prod", "similarity") or save_embeddings
        P_e, a = scorer.project_sequences(P_f, mode, want_embedding=need_emb)
        L_e, c = self._projected_labels(scorer, L_f, mode, need_emb)
        embeddings = {"output_layer_embeddings": [], "joint_embeddings": []}
        if self.feature_fusion == "similarity":
            logits = scorer.score_similarity(P_e, L_e, self.temperature, mode)
        elif not save_embeddings:
            logits = scorer.score(a, c, P_e, L_e, mode)
        else:
            # Diagnostics contract of the reference (ProtNote.py:294-303,324-334): the joint features and the last hidden
            # layer for every pair, on the CPU.  The hidden layer comes out of the scorer kernel's epilogue; the joint
            # tensor is a pure gather of P_e / L_e rows (it is only materialised here, never on the scoring path).
            B, L = P_e.shape[0], L_e.shape[0]
            hidden = torch.empty(B * L, self._cfg["out_hidden"], dtype=torch.float32, device=dev)
            logits = scorer.score(a, c, P_e, L_e, mode, hidden_out=hidden)
            p = P_e[:, None, :].expand(B, L, P_e.shape[1])
            t = L_e[None, :, :].expand(B, L, L_e.shape[1])
            parts = [p, t]
            if self.feature_fusion == "concatenation_diff":
                parts.append(p - t)
            elif self.feature_fusion == "concatenation_prod":
                parts.append(p * t)
            embeddings["joint_embeddings"] = torch.cat(parts, dim=2).reshape(B * L, -1).detach().cpu()
            embeddings["output_layer_embeddings"] = hidden.cpu()
        if self.match_autocast_dtype and torch.is_autocast_enabled():
            logits = logits.to(torch.get_autocast_gpu_dtype())
        return logits, embeddings
