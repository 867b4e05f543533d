"""world_size-2 gloo test of the label-sharded TRAINING step: every rank owns a contiguous block of label rows, the
BatchNorm sums (forward) and the two BatchNorm-backward sums are all-reduced, and the summed per-rank gradients must
equal the single-process gradients of the whole B x L batch.  The per-rank compute step is the
torch stand-in for the CUDA primitives (oracle/train_ops.py); the sequencing under test is protnote_b200/train.py."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.cases import CASES
from oracle.protnote_oracle import synth_state_dict
from oracle.train_ops import TorchOps
from oracle.train_oracle import synth_targets, train_step_oracle
from protnote_b200 import train as pn_train
from protnote_b200.sharded import label_row_bounds
from tests.helpers import build_b200_model


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem(variant=False):
    ecfg, scfg, *_ = CASES["tiny_concat"]
    import dataclasses
    if variant == "prod":        # [p; t; p * t]: layer 1 is an ordinary (statistics-all-reduced) layer plus two marginals
        scfg = dataclasses.replace(scfg, feature_fusion="concatenation_prod")
    elif variant == "dropout":   # OUTPUT_MLP_DROPOUT: one base seed for all ranks, rank-salted masks on the sharded rows
        scfg = dataclasses.replace(scfg, output_mlp_dropout=0.3)
    elif variant:    # output MLP without BatchNorm (hidden biases, nothing to all-reduce in its backward) on [p; t; p - t]
        scfg = dataclasses.replace(scfg, output_mlp_batchnorm=False, feature_fusion="concatenation_diff")
    sd = synth_state_dict(ecfg, scfg, seed=42, calib_T=64)
    g = torch.Generator().manual_seed(21)
    B, L = 5, 11     # 11 rows over 2 ranks: uneven shards
    return ecfg, scfg, sd, torch.randn(B, 72, generator=g), torch.randn(L, 40, generator=g), synth_targets(B, L, 21)


def _worker(rank, world, port, q, fused=False, variant=False):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.set_num_threads(2)
        ecfg, scfg, sd, P_f, L_f, y = _problem(variant)
        model = build_b200_model(ecfg, scfg, sd, device="cpu").double().train()
        comm = pn_train.Comm()
        ls, le = label_row_bounds(L_f.shape[0], 1, rank, world)
        if variant == "dropout":
            torch.manual_seed(1000 + rank)       # ranks seeded differently, as ProtNoteTrainer's seed_everything(seed + rank)
        if fused:
            # loss fused into the last forward primitive (focal, the reference's default LOSS_FN) and the parameter
            # gradients all-reduced inside the backward, overlapped with it: no allreduce_gradients call
            loss, logits = pn_train.train_loss(model, P_f.double(), L_f[ls:le].double(), y[:, ls:le].double(), loss="focal",
                                               gamma=2.0, alpha=0.25, ops=TorchOps(torch.float64), comm=comm,
                                               L_total=L_f.shape[0], reduce_gradients=True)
            loss.backward()
        else:
            logits = pn_train.train_logits(model, P_f.double(), L_f[ls:le].double(), ops=TorchOps(torch.float64), comm=comm,
                                           L_total=L_f.shape[0])
            # the loss of the whole batch is the mean over all B x L pairs: each rank contributes its slab's sum
            loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, y[:, ls:le].double(), reduction="sum") / y.numel()
            loss.backward()
            pn_train.allreduce_gradients(model, comm)
        total = loss.detach().clone()
        dist.all_reduce(total)
        if rank == 0:
            # numpy copies: pickled by value (torch tensors would travel as shared-memory handles that die with the rank)
            out = {k: p.grad.numpy().copy() for k, p in model.named_parameters() if p.grad is not None}
            bufs = {k: b.numpy().copy() for k, b in model.named_buffers()
                    if "running_" in k and not k.startswith("sequence_encoder")}
            q.put((0, None, float(total), out, bufs, logits.detach().numpy().copy()))
        else:
            q.put((rank, None))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
        raise


import pytest  # noqa: E402


def _sharded_masks(ecfg, scfg, sd, B, L, world):
    """The masks of the label-sharded step as one set of global multipliers: the base seed is rank 0's first CPU draw
    after torch.manual_seed(1000) (broadcast to all ranks); W_p's sites are the same on every rank, W_l's and the output
    MLP's are each rank's own rows (rank-salted seeds, local row indices) and are stitched together along the label axis."""
    from oracle.train_ops import dropout_multiplier
    torch.manual_seed(1000)
    base = int(torch.randint(0, 1 << 62, (1,), dtype=torch.int64))
    model = build_b200_model(ecfg, scfg, sd, device="cpu")
    masks = {}
    plans = [pn_train.dropout_sites(model, base, r) for r in range(world)]
    bounds = [label_row_bounds(L, 1, r, world) for r in range(world)]
    for (t, i), (seed0, p, width) in plans[0].items():
        if t == "p":
            assert all(pl[(t, i)][0] == seed0 for pl in plans)
            masks[(t, i)] = dropout_multiplier(seed0, B, width, p)
        elif t == "l":
            assert len({pl[(t, i)][0] for pl in plans}) == world
            masks[(t, i)] = torch.cat([dropout_multiplier(pl[(t, i)][0], le - ls, width, p)
                                       for pl, (ls, le) in zip(plans, bounds)])
        else:
            masks[(t, i)] = torch.cat([dropout_multiplier(pl[(t, i)][0], B * (le - ls), width, p).reshape(B, le - ls, width)
                                       for pl, (ls, le) in zip(plans, bounds)], dim=1).reshape(B * L, width)
    return masks


@pytest.mark.parametrize("fused,variant", [(False, False), (True, False), (True, True), (True, "dropout"), (True, "prod")],
                         ids=["bce_via_autograd", "fused_focal_overlapped_allreduce", "fused_focal_diff_no_batchnorm",
                              "fused_focal_output_mlp_dropout", "fused_focal_prod"])
def test_label_sharded_training_step_equals_single_process(fused, variant):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, fused, variant)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for r in results:
        assert r[1] is None, r
    r0 = next(r for r in results if r[0] == 0)
    _, _, loss, grads, bufs, logits0 = r0
    grads = {k: torch.from_numpy(v) for k, v in grads.items()}
    bufs = {k: torch.from_numpy(v) for k, v in bufs.items()}
    logits0 = torch.from_numpy(logits0)
    ecfg, scfg, sd, P_f, L_f, y = _problem(variant)
    kw = dict(loss="focal", gamma=2.0, alpha=0.25) if fused else {}
    if variant == "dropout":
        kw["masks"] = _sharded_masks(ecfg, scfg, sd, P_f.shape[0], L_f.shape[0], world)
    o_logits, o_loss, o_grads, o_stats = train_step_oracle(sd, P_f, L_f, y, scfg, **kw)
    ls, le = label_row_bounds(L_f.shape[0], 1, 0, world)
    assert (logits0 - o_logits[:, ls:le]).abs().max() < 1e-9
    assert abs(loss - float(o_loss)) < (1e-7 if fused else 1e-10)      # the fused loss leaves each rank as fp32
    for k, g in o_grads.items():
        assert (grads[k] - g).abs().max() <= 1e-9 * max(1.0, float(g.abs().max())), k
    for k, v in o_stats.items():
        assert (bufs[k] - v).abs().max() < 1e-9, k
