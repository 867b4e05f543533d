"""Label-sharded multi-GPU scoring (one process per GPU, torch.distributed).

Every (protein, label) pair is independent in eval mode, so the path shards without any exchange inside the
arithmetic.  Following BASELINE.json's north_star the LABEL axis is partitioned (contiguous blocks of whole labels,
i.e. groups of k description rows stay on one rank) and the per-rank logit slabs [B, L/(k*W)] are all-gathered once;
to avoid W-fold redundant encoder work the proteins are sharded for the encoder and the pooled embeddings
[B, C] (18 MB at B=4096) are all-gathered first.  Both collectives are plain NCCL all-gathers over NVLink: there is
no compute step that a transfer could be fused into (the payload is 4 bytes per 38 MFLOP pair).

The reference has no counterpart: it shards SEQUENCES with DistributedDataParallel (bin/main.py:452) and never
gathers logits; its DISTRIBUTE_LABELS knob is vestigial (protnote/data/collators.py:81-91, samplers.py:236-238).

The functions take the two compute steps as callables so that the partition / gather logic is testable on CPU with
the gloo backend (tests/test_sharded_gloo.py drives it with the oracle as the compute step).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [start, end) of `n_items` owned by `rank`; the first n_items % world ranks get one more."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def label_row_bounds(n_rows: int, k: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [start, end) of the label-embedding matrix owned by `rank`; the k rows of one label never split."""
    if n_rows % k != 0:
        raise ValueError(f"{n_rows} label rows is not a multiple of descriptions_per_label={k}")
    s, e = shard_bounds(n_rows // k, rank, world)
    return s * k, e * k


def all_gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather of row blocks with the shard_bounds partition: local [n_local, D] -> [n_total, D] on every rank."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = (n_total + world - 1) // world
    padded = local.new_zeros((per,) + tuple(local.shape[1:]))
    padded[: local.shape[0]] = local
    out = local.new_empty((world * per,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    pieces = []
    for r in range(world):
        s, e = shard_bounds(n_total, r, world)
        pieces.append(out[r * per: r * per + (e - s)])
    del rank
    return torch.cat(pieces, 0)


def all_gather_columns(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather of column blocks (the logit slab): local [B, n_local] -> [B, n_total] on every rank."""
    world = dist.get_world_size(group)
    B = local.shape[0]
    per = (n_total + world - 1) // world
    padded = local.new_zeros((B, per))
    padded[:, : local.shape[1]] = local
    flat = local.new_empty((world * B, per))      # rank-major concatenation along dim 0 (what gloo and NCCL both accept)
    dist.all_gather_into_tensor(flat, padded.contiguous(), group=group)
    out = flat.view(world, B, per)
    pieces = []
    for r in range(world):
        s, e = shard_bounds(n_total, r, world)
        pieces.append(out[r, :, : e - s])
    return torch.cat(pieces, 1)


def sharded_forward(encode: Callable[[torch.Tensor, torch.Tensor], torch.Tensor],
                    score: Callable[[torch.Tensor, torch.Tensor], torch.Tensor],
                    sequence_onehots: torch.Tensor, sequence_lengths: torch.Tensor, label_embeddings: torch.Tensor,
                    descriptions_per_label: int = 1, group=None, inputs_are_local: bool = False,
                    total_sequences: Optional[int] = None, total_label_rows: Optional[int] = None) -> torch.Tensor:
    """logits [B, L/k] on every rank.

    encode(onehots [b, Cin, T], lengths [b]) -> P_f [b, C];  score(P_f [B, C], label_rows [l, D]) -> [B, l/k].
    With inputs_are_local=False every rank is handed the full inputs and slices its shard; with True the caller
    already holds only this rank's proteins / label rows (shard_bounds / label_row_bounds partition) and passes the totals.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    k = descriptions_per_label
    if world == 1:
        return score(encode(sequence_onehots, sequence_lengths), label_embeddings)
    if inputs_are_local:
        B, L = int(total_sequences), int(total_label_rows)
        x, lens, lab = sequence_onehots, sequence_lengths, label_embeddings
    else:
        B, L = sequence_onehots.shape[0], label_embeddings.shape[0]
        ps, pe = shard_bounds(B, rank, world)
        ls, le = label_row_bounds(L, k, rank, world)
        x, lens, lab = sequence_onehots[ps:pe], sequence_lengths[ps:pe], label_embeddings[ls:le]
    P_local = encode(x, lens)
    P_f = all_gather_rows(P_local, B, group)
    logits_local = score(P_f, lab)
    return all_gather_columns(logits_local, L // k, group)


def native_sharded_forward(model, sequence_onehots, sequence_lengths, label_embeddings, group=None, **kw):
    """sharded_forward with the sm_100a model (protnote_b200.ProtNote.ProtNote in eval mode) as the compute step.

    For the concatenation fusions the PROTEIN-side head is sharded with the encoder: each rank runs W_p and the protein half
    of output layer 1 on its own proteins only and the [B, H] halves a[b] are what is all-gathered (the pair scorer needs
    nothing else from the protein side) - in strict mode that head runs in fp64 on the CUDA cores (47 ms for 4096 proteins),
    which replicated on every rank was 2.7 % of the 8-GPU step."""
    from . import native
    fusion = model.feature_fusion
    if model.training or not fusion.startswith("concatenation") or fusion == "concatenation_prod":
        def encode(x, lens):
            return model.sequence_encoder.get_embeddings(x, lens)

        def score(P_f, lab):
            return model(sequence_embeddings=P_f, label_embeddings=lab)[0]

        return sharded_forward(encode, score, sequence_onehots, sequence_lengths, label_embeddings,
                               model.inference_descriptions_per_label, group, **kw)
    mode = native.MODES[model.precision]
    dev = next(model.W_p.parameters()).device

    def encode(x, lens):        # -> this rank's rows of a [b, H]
        scorer = model._ensure_packed()
        P_f = model.sequence_encoder.get_embeddings(x, lens)
        return scorer.project_sequences(P_f, mode, want_embedding=False)[1]

    def score(a, lab):          # a [B, H] (all proteins), lab = this rank's label rows
        scorer = model._ensure_packed()
        _, c = model._projected_labels(scorer, lab.to(dev, non_blocking=True), mode, False)
        return scorer.score(a, c, None, None, mode)

    return sharded_forward(encode, score, sequence_onehots, sequence_lengths, label_embeddings,
                           model.inference_descriptions_per_label, group, **kw)
