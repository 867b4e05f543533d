"""world_size-2 gloo test of the label-sharded forward (partition + the two all-gathers), with the CPU oracle as the
per-rank compute step: sharded result == single-process result, bitwise (sharding the label axis does not change any
reduction inside a pair)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.protnote_oracle import proteinfer_embeddings, score_pairs
from protnote_b200.sharded import label_row_bounds, shard_bounds, sharded_forward
from tests.helpers import load_case


def test_shard_bounds_cover_everything():
    for n in (0, 1, 5, 8, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert label_row_bounds(20, 2, 1, 3) == (8, 14)      # labels 4..7 -> rows 8..14: the k rows of a label stay together
    with pytest.raises(ValueError):
        label_row_bounds(21, 2, 0, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, inputs_are_local, q):
    try:
        _worker_body(rank, world, port, case, inputs_are_local, q)
    except Exception as e:  # noqa: BLE001 - report instead of leaving the parent waiting on the queue
        q.put((rank, repr(e)))
        raise


def _worker_body(rank, world, port, case, inputs_are_local, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    ecfg, scfg, sd, onehots, lengths, labels, g = load_case(case)
    k = scfg.inference_descriptions_per_label

    def encode(x, lens):
        return proteinfer_embeddings(sd, x, lens, ecfg, "sequence_encoder.")

    def score(P_f, lab):
        return score_pairs(sd, P_f, lab, scfg)

    with torch.no_grad():
        if inputs_are_local:
            ps, pe = shard_bounds(onehots.shape[0], rank, world)
            ls, le = label_row_bounds(labels.shape[0], k, rank, world)
            out = sharded_forward(encode, score, onehots[ps:pe], lengths[ps:pe], labels[ls:le], k,
                                  inputs_are_local=True, total_sequences=onehots.shape[0],
                                  total_label_rows=labels.shape[0])
        else:
            out = sharded_forward(encode, score, onehots, lengths, labels, k)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case,inputs_are_local", [("tiny_concat", False), ("tiny_k2", True)])
def test_label_sharded_equals_single_process(case, inputs_are_local):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, inputs_are_local, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=240) for _ in range(world))
    for r, v in results.items():
        assert isinstance(v, torch.Tensor), f"rank {r} failed: {v}"
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ecfg, scfg, sd, onehots, lengths, labels, g = load_case(case)
    with torch.no_grad():
        single = score_pairs(sd, proteinfer_embeddings(sd, onehots, lengths, ecfg, "sequence_encoder."), labels, scfg)
    for r in range(world):
        assert results[r].shape == g["logits"].shape
        # per-pair arithmetic is identical; only the GEMM blocking of the label axis differs between 1 and 2 shards
        assert (results[r] - single).abs().max().item() <= 2e-5
    assert torch.equal(results[0], results[1])
