"""The drop-in wrapped the way bin/main.py:452 wraps the reference: DistributedDataParallel(find_unused_parameters=True)
over NCCL, `.module` reachable (ProtNoteTrainer.py:194-197), forward called with keywords only.  One golden case in eval mode
and one training step (gradients through DDP's reducer) on a single-rank NCCL group; the label-sharded training step on
two ranks (needs two GPUs) against the single-process CPU training oracle."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from tests.helpers import build_b200_model, load_case, topk_agree

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.fixture()
def nccl_single_rank():
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(_free_port())
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    yield dist
    dist.destroy_process_group()


def test_ddp_wrapped_eval_and_training_step(nccl_single_rank):
    from torch.nn.parallel import DistributedDataParallel as DDP
    from oracle.make_golden_train import train_inputs
    from oracle.train_oracle import train_step_oracle
    # ---- eval: golden case through the DDP wrapper, as ProtNoteTrainer.evaluation_step calls it (under autocast)
    ecfg, scfg, sd, onehots, lengths, labels, g = load_case("tiny_concat")
    model = build_b200_model(ecfg, scfg, sd, device="cuda:0")
    ddp = DDP(model, device_ids=[0], find_unused_parameters=True)
    assert ddp.module is model
    ddp.eval()
    with torch.no_grad(), torch.autocast("cuda"):
        logits, extra = ddp(sequence_onehots=onehots.cuda(), sequence_lengths=lengths.cuda(), label_embeddings=labels.cuda())
    assert logits.dtype == torch.float32          # fp32-grade logits even under autocast
    assert (logits.cpu() - g["logits"]).abs().max() <= 1e-4
    assert topk_agree(g["logits"], logits.cpu(), k=10, tol=1e-4)
    model.match_autocast_dtype = True             # opt-in dtype drop-in: what the reference returns under autocast
    with torch.no_grad(), torch.autocast("cuda"):
        logits16, _ = ddp(sequence_onehots=onehots.cuda(), sequence_lengths=lengths.cuda(), label_embeddings=labels.cuda())
    assert logits16.dtype == torch.float16 and torch.equal(logits16, logits.half())
    # ---- one training step through DDP's gradient reducer, checked against the training oracle
    ecfg, scfg, sd, P_f, L_f, y = train_inputs("train_tiny")
    tmodel = build_b200_model(ecfg, scfg, sd, device="cuda:0")
    for p in tmodel.sequence_encoder.parameters():     # frozen, as main_utils / ProtNoteTrainer set it up
        p.requires_grad_(False)
    tddp = DDP(tmodel, device_ids=[0], find_unused_parameters=True)
    tddp.train()
    tlogits, _ = tddp(sequence_embeddings=P_f.cuda(), label_embeddings=L_f.cuda())
    loss = torch.nn.functional.binary_cross_entropy_with_logits(tlogits, y.cuda())
    loss.backward()
    torch.cuda.synchronize()
    o_logits, o_loss, o_grads, _ = train_step_oracle(sd, P_f, L_f, y, scfg)
    assert (tlogits.detach().cpu().double() - o_logits).abs().max() <= 1e-4
    assert abs(float(loss.detach()) - float(o_loss)) <= 1e-5
    named = dict(tddp.module.named_parameters())
    for k, og in o_grads.items():
        got = named[k].grad
        assert got is not None, k
        assert float((got.cpu().double() - og).norm() / og.norm().clamp_min(1e-30)) <= 1e-3, k


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_label_sharded_training_step_on_two_nccl_ranks():
    """bench.py's train_parity_small on a real 2-rank NCCL group: logits, loss and every gradient of the label-sharded step
    equal the single-process CPU oracle's (1e-4 on logits, 1e-3 of each gradient's largest entry)."""
    code = ("import json, os, torch, torch.distributed as dist, bench\n"
            "r, w, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])\n"
            "torch.cuda.set_device(lr)\n"
            "import datetime\n"
            "dist.init_process_group('nccl', device_id=torch.device('cuda', lr), timeout=datetime.timedelta(seconds=90))\n"
            "out = bench.train_parity_small(torch.device('cuda', lr), r, w)\n"
            "if r == 0: print('PARITY ' + json.dumps(out), flush=True)\n"
            # more ranks than sequences: rank 1 owns none and must still take part in every BatchNorm-sum all-reduce
            "from tests.helpers import build_b200_model, load_case\n"
            "from protnote_b200.sharded import all_gather_rows, shard_bounds\n"
            "ecfg, scfg, sd, onehots, lengths, _, _ = load_case('tiny_concat')\n"
            "enc = build_b200_model(ecfg, scfg, sd, device=torch.device('cuda', lr)).train().sequence_encoder\n"
            "enc.train_shard = (None, 1)\n"
            "ps, pe = shard_bounds(1, r, w)\n"
            "with torch.no_grad():\n"
            "    e = all_gather_rows(enc.get_embeddings(onehots[:1][ps:pe].cuda(), lengths[:1][ps:pe].cuda()), 1)\n"
            "torch.cuda.synchronize()\n"
            "if r == 0: print('ZERO_RANK_OK ' + str(tuple(e.shape)), flush=True)\n"
            "dist.destroy_process_group()\n")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), "--no-python",
                          sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert res.returncode == 0, res.stderr[-3000:]
    import json
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("PARITY ")][-1]
    out = json.loads(line[len("PARITY "):])
    assert out["ranks"] == 2 and out["within_tol"], out
    assert out["max_abs_logit_err"] <= 1e-4 and out["abs_loss_err"] <= 1e-5
    assert out["worst_gradient_max_err_over_max_entry"] <= 1e-3
    assert "ZERO_RANK_OK (1, 72)" in res.stdout
