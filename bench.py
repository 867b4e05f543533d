#!/usr/bin/env python
"""Benchmark of the ProtNote scoring hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode strict|fast]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path (encoder -> mean pool -> W_p / W_l -> fused pair scorer) over the configuration the
metric is quoted on: 4096 synthetic 1024-residue sequences x 32768 label-embedding rows, fp32 in / fp32 out
(BASELINE.json configs[1]).  Rank 0 prints ONE JSON line.

  value     pair-scores/s with every input already resident in HBM (CUDA events, max over ranks)
  e2e       the same metric through the module's public forward() with HOST (pinned) tensors: H2D of the one-hot
            sequences, lengths and label embeddings and D2H of the logits happen inside the timed region
  roofline  the pair scorer's GEMM kernel (tcgen05): algorithmic FLOPs / CUDA-event launch time vs the measured
            bf16 tensor peak in MEASURED_PEAKS.json
  cpu_baseline  the oracle (CPU restatement of the reference forward, oracle/) timed on the host cores, bounded sample
  --impl reference   times that CPU implementation alone (rank 0), same metric / unit
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "(protein,label) pair-scores/sec at 1024 aa x 32K labels"
UNIT = "pair-scores/s"
FLOP_PER_PAIR = 37_754_880          # 2*(2*3072^2) + 2*3072  
B_TOTAL, T_LEN, L_ROWS = 4096, 1024, 32768
CPU_SAMPLE = (8, 1024, 4096)        # sequences, residues, label rows of the bounded CPU sample


def config_tag(B, T, L, k):
    if (B, T, L, k) == (B_TOTAL, T_LEN, L_ROWS, 1):
        return "BASELINE.json configs[1]"
    if (B, L, k) == (10000, 10268, 2):
        return "BASELINE.json configs[3]: zero-shot EC shape, 5134 EC numbers x 2 descriptions"
    return "non-headline shape"


def base_config_model(precision: str, descriptions_per_label: int = 1):
    """Random-init ProtNote with the published architecture (configs/base_config.yaml)."""
    from protnote_b200.ProtNote import ProtNote
    from protnote_b200.protein_encoders import ProteInfer
    torch.manual_seed(42)
    enc = ProteInfer(num_labels=8, input_channels=20, output_channels=1100, kernel_size=9, activation=torch.nn.ReLU,
                     dilation_base=3, num_resnet_blocks=5, bottleneck_factor=0.5, precision=precision)
    model = ProtNote(protein_embedding_dim=1100, label_embedding_dim=1024, latent_dim=1024,
                     label_embedding_pooling_method="mean", sequence_encoder=enc,
                     inference_descriptions_per_label=descriptions_per_label,
                     output_mlp_hidden_dim_scale_factor=3, output_mlp_num_layers=3, outout_mlp_add_batchnorm=True,
                     projection_head_num_layers=4, projection_head_hidden_dim_scale_factor=3,
                     feature_fusion="concatenation", precision=precision)
    g = torch.Generator().manual_seed(7)
    for m in model.modules():       # non-trivial BatchNorm statistics so that the folding is exercised
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)
    return model.eval()


def synthetic_inputs(B, T, L, pinned):
    g = torch.Generator().manual_seed(1234)
    tokens = torch.randint(0, 20, (B, T), generator=g)
    onehots = torch.zeros(B, 20, T)
    onehots.scatter_(1, tokens[:, None, :], 1.0)
    lengths = torch.full((B,), T, dtype=torch.long)
    labels = torch.randn(L, 1024, generator=torch.Generator().manual_seed(4321))
    if pinned:
        onehots, lengths, labels = onehots.pin_memory(), lengths.pin_memory(), labels.pin_memory()
    return onehots, lengths, labels


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001 - nvidia-smi missing: report no clocks
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = sorted(sm)[len(sm) // 4:] if sm else []
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, "fallback (B200_PROFILING.md sustained figure)"


def cpu_oracle_rate(reps: int, warmup: int):
    """pair-scores/s of the CPU restatement of the reference forward (oracle/) on the host cores."""
    from oracle.protnote_oracle import EncoderCfg, ScorerCfg, protnote_forward, synth_inputs
    B, T, L = CPU_SAMPLE
    ecfg, scfg = EncoderCfg(), ScorerCfg()
    model = base_config_model("strict")
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    onehots, lengths, labels = synth_inputs(B, T, L, ecfg, scfg, ragged=False, seed=1234)
    cores = torch.get_num_threads()
    for _ in range(warmup):
        protnote_forward(sd, onehots, lengths, labels, ecfg, scfg)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        protnote_forward(sd, onehots, lengths, labels, ecfg, scfg)
        times.append(time.perf_counter() - t0)
    mean = sum(times) / len(times)
    sample = (f"{B} seqs x {T} aa x {L} label rows per step, fp32, torch {torch.__version__} CPU ops "
              f"(oracle/protnote_oracle.py), {reps} timed reps")
    return B * L / mean, mean * 1e3, cores, sample


def run_reference(args, rank, world):
    if rank != 0:
        return
    rate, ms, cores, sample = cpu_oracle_rate(reps=max(1, args.steps), warmup=max(1, min(args.warmup, 1)))
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"inference {B_TOTAL} x {T_LEN} aa x {L_ROWS} label rows, fp32 in/out (BASELINE.json configs[1])",
                       "mode": "reference arithmetic, fp32, CPU", "sequences": B_TOTAL, "seq_len": T_LEN, "label_rows": L_ROWS,
                       "descriptions_per_label": 1, "parallelism": "host cores of rank 0",
                       "sample": "each step is a bounded sample of the workload: " + sample},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from protnote_b200 import native
    from protnote_b200.sharded import label_row_bounds, native_sharded_forward, shard_bounds

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    B, T, L = args.sequences, args.seq_len, args.labels
    kdesc = args.descriptions_per_label
    model = base_config_model(args.mode, kdesc).to(dev)
    onehots_h, lengths_h, labels_h = synthetic_inputs(B, T, L, pinned=True)
    ps, pe = shard_bounds(B, rank, world)
    ls, le = label_row_bounds(L, kdesc, rank, world)
    # this rank's shard of the inputs (world == 1: everything)
    x_h, len_h, lab_h = onehots_h[ps:pe], lengths_h[ps:pe], labels_h[ls:le]
    if world > 1:
        x_h, len_h, lab_h = x_h.contiguous().pin_memory(), len_h.contiguous().pin_memory(), lab_h.contiguous().pin_memory()
    x_d, len_d, lab_d = x_h.to(dev), len_h.to(dev), lab_h.to(dev)
    logits_host = torch.empty(B, (le - ls) // kdesc, dtype=torch.float32).pin_memory()

    def forward(x, lens, lab):
        model._label_cache = None          # W_l(label_embeddings) is recomputed every step, like the reference does
        with torch.no_grad():
            if world == 1:
                return model(sequence_onehots=x, sequence_lengths=lens, label_embeddings=lab)[0]
            return native_sharded_forward(model, x, lens, lab, inputs_are_local=True, total_sequences=B,
                                          total_label_rows=L)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    out = None

    def step_device():
        nonlocal out
        out = forward(x_d, len_d, lab_d)

    def step_e2e():
        nonlocal out
        x = x_h.to(dev, non_blocking=True)
        lens = len_h.to(dev, non_blocking=True)
        lab = lab_h.to(dev, non_blocking=True)
        out = forward(x, lens, lab)
        logits_host.copy_(out[:, ls // kdesc:le // kdesc] if world > 1 else out, non_blocking=True)

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = native.launch_count()
    native.gemm_timing(True)
    ms_step = timed(step_device, args.steps)
    gemm_ms, gemm_launches, gemm_flops = native.gemm_timing_read()
    native.gemm_timing(False)
    launches = native.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    finite = bool(torch.isfinite(out).all().item())

    step_e2e()                              # one untimed e2e step (pinned-buffer / allocator warm-up)
    ms_e2e = timed(step_e2e, args.steps)

    # per-stage device times of one step on this rank's shard (outside the timed regions; explains `value`)
    breakdown = {}
    with torch.no_grad():
        scorer = model._ensure_packed()
        mode = native.MODES[args.mode]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        torch.cuda.synchronize()
        ev[0].record()
        P_f = model.sequence_encoder.get_embeddings(x_d, len_d)
        ev[1].record()
        _, a = scorer.project_sequences(P_f, mode)
        ev[2].record()
        _, c = scorer.project_labels(lab_d, mode)
        ev[3].record()
        scorer.score(a, c, mode=mode)
        ev[4].record()
        torch.cuda.synchronize()
        for i, name in enumerate(("encoder", "W_p+layer1_p", "W_l+layer1_l", "pair_scorer")):
            breakdown[name + "_ms"] = ev[i].elapsed_time(ev[i + 1])
        breakdown["encoder_residues_per_s"] = x_d.shape[0] * T / (breakdown["encoder_ms"] * 1e-3)
        breakdown["encoder_algorithmic_tflops"] = x_d.shape[0] * T * 60_896_000 / (breakdown["encoder_ms"] * 1e-3) / 1e12
        breakdown["scorer_pairs_per_s"] = a.shape[0] * c.shape[0] / (breakdown["pair_scorer_ms"] * 1e-3)

    # the same step with one tensor-core pass per product (fp16 operands, fp32 accumulate = the arithmetic the reference
    # itself uses on a GPU under torch.autocast): how fast the kernels are when fp32-grade logits are not required, and
    # how far those logits are from the strict ones
    fast = None
    if args.mode == "strict" and world == 1 and not args.no_fast:
        strict_out = out
        model.precision = model.sequence_encoder.precision = "fast"
        step_device()
        native.gemm_timing(True)
        ms_fast = timed(step_device, 1)
        f_ms, f_n, f_flops = native.gemm_timing_read()
        native.gemm_timing(False)
        model.precision = model.sequence_encoder.precision = "strict"
        diff = (out - strict_out).abs()
        k10 = min(10, out.shape[1])
        same = (out.topk(k10, dim=1).indices == strict_out.topk(k10, dim=1).indices).all(dim=1).float().mean()
        fast = {"value": B * L / (ms_fast * 1e-3), "unit": UNIT, "ms_per_step": ms_fast,
                "scorer_gemm_tflops": f_flops / (f_ms * 1e-3) / 1e12 if f_ms > 0 else None,
                "max_abs_logit_diff_vs_strict": float(diff.max()), "mean_abs_logit_diff_vs_strict": float(diff.mean()),
                "logit_std": float(strict_out.std()), "top10_identical_fraction": float(same),
                "max_abs_logit_diff_over_logit_std": float(diff.max() / strict_out.std().clamp_min(1e-30)),
                "note": "NOT the headline: fp16-operand arithmetic is a few percent of the logit std away from the fp32 "
                        "result (5e-2 at logit std 2 on the calibrated golden cases), far outside the 1e-4 bar; this random-init "
                        "bench model has un-calibrated BatchNorm statistics, hence the small logit std"}
        out = strict_out
    if rank != 0:
        return
    pairs = B * L
    peak, peak_src = measured_peaks()
    passes = 3 if args.mode == "strict" else 1
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    # whole-job bytes per step: every rank copies its own protein / label shard in and its own logit slab out
    h2d = onehots_h.numel() * 4 + lengths_h.numel() * 8 + labels_h.numel() * 4
    d2h = B * (L // kdesc) * 4
    line = {
        "metric": METRIC, "value": pairs / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32 (fp16 hi/lo planes, 3 tcgen05 passes, fp32 accumulate)" if passes == 3
        else "f16 operands, fp32 accumulate", "data": "synthetic",
        "config": {"workload": f"inference {B} x {T} aa x {L} label rows, fp32 in/out ({config_tag(B, T, L, kdesc)})",
                   "mode": args.mode, "sequences": B, "seq_len": T, "label_rows": L, "descriptions_per_label": kdesc,
                   "parallelism": "1 GPU" if world == 1 else f"label-sharded x{world} (proteins sharded for the encoder), "
                                                              "NCCL all-gather of P_f and of the logit slab",
                   "l2": "inputs (one-hots 335 MB + label embeddings 134 MB) are larger than the 126 MB L2",
                   "label_projection": "recomputed every step (cache cleared)"},
        "e2e": {"value": pairs / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": "pn::gemm_kernel<32,3> (pair-scorer output-MLP layers 2 and 3)" if passes == 3
                     else "pn::gemm_kernel<64,1>", "bound": "tensor", "achieved": achieved, "peak": peak,
                     "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full`
                     # capture (profiles/r01_ncu_full_gemm_scorer_encoder.txt): 31.5 KB/row for the layer that stores its
                     # activations, 13.2 KB/row for the dot-epilogue layer -> 22.4 KB per row per launch on average;
                     # algorithmic: 12 KB/row of A planes read (+12 KB/row written by the storing layer) = 18 KB/row
                     "traffic": (22.4e3 * gemm_flops / (2.0 * 3072 * 3072) / gemm_launches) if gemm_launches else None,
                     "traffic_unit": "bytes per launch (ncu dram bytes, scaled by the rows of the timed launches)",
                     "algorithmic_bytes_per_launch": (18.0e3 * gemm_flops / (2.0 * 3072 * 3072) / gemm_launches)
                     if gemm_launches else None,
                     "peak_source": peak_src, "launches": int(gemm_launches),
                     "avg_launch_ms": gemm_ms / gemm_launches if gemm_launches else None,
                     "share_of_step": gemm_ms / (ms_step * args.steps) if ms_step > 0 else None,
                     "tensor_passes": passes, "executed_tflops": achieved * passes,
                     "executed_frac": achieved * passes / peak if peak else None,
                     "note": "achieved counts ALGORITHMIC flops (2*M*N*K); strict mode executes 3 fp16 passes per "
                             "algorithmic flop to reach fp32 accuracy, so its ceiling is peak/3"},
        "roofline_encoder": {"kernel": "pn::gemm_kernel (11 convolutions of the ProteInfer encoder as implicit GEMMs)",
                             "bound": "tensor", "achieved": breakdown["encoder_algorithmic_tflops"], "peak": peak,
                             "unit": "TFLOP/s", "frac": breakdown["encoder_algorithmic_tflops"] / peak if peak else None,
                             "executed_frac": breakdown["encoder_algorithmic_tflops"] * passes / peak if peak else None,
                             "share_of_step": breakdown["encoder_ms"] / ms_step},
        "breakdown_rank0": breakdown,
        "outputs_finite": finite,
    }
    if fast is not None:
        fast["frac_of_peak_scorer_gemm"] = fast["scorer_gemm_tflops"] / peak if fast["scorer_gemm_tflops"] else None
        line["fast_mode"] = fast
    if world == 1 and not args.no_cpu_baseline:
        rate, ms, cores, sample = cpu_oracle_rate(reps=2, warmup=1)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                "ms_per_sample": ms}
    print(json.dumps(line), flush=True)


TRAIN_B = 64
# algorithmic FLOPs of one training step per (protein, label-row) pair: forward 2 GEMMs + dot, backward dgrad + wgrad of
# the same two layers (layer 1 is factorised in both directions) -> 3 x the forward GEMM work
TRAIN_FLOP_PER_PAIR = 3 * 2 * (2 * 3072 * 3072) + 2 * 3072
TRAIN_FLOP_PER_LABEL_ROW = 3 * 50_331_648 + 3 * 2 * 1024 * 3072     # W_l forward + backward, label half of layer 1
TRAIN_FLOP_PER_PROTEIN = 3 * 50_798_592 + 3 * 2 * 1024 * 3072 + 1024 * 60_896_000   # W_p, protein half, frozen encoder


def run_train(args, rank, world, local_rank):
    """BASELINE.json configs[2]: one training step = frozen encoder forward -> W_p / W_l / output MLP forward with BATCH
    statistics -> BCE-with-logits -> backward -> gradient all-reduce (label-sharded ranks) -> Adam.  Label rows are
    sharded over ranks; BatchNorm sums are all-reduced so the step equals the single-process step on the whole batch."""
    import torch.distributed as dist
    from protnote_b200 import native, train as pn_train
    from protnote_b200.sharded import label_row_bounds

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    B = TRAIN_B if args.sequences == B_TOTAL else args.sequences
    T, L = args.seq_len, args.labels
    mode = args.mode if args.train_mode is None else args.train_mode
    # model.train() as ProtNoteTrainer.train does it: the frozen encoder's BatchNorm layers use batch statistics too
    model = base_config_model(mode).to(dev).train()
    params = pn_train.trainable_parameters(model)
    for p in model.sequence_encoder.parameters():
        p.requires_grad_(False)
    opt = torch.optim.Adam(params, lr=3e-4, fused=True)
    comm = pn_train.Comm() if world > 1 else None
    onehots_h, lengths_h, labels_h = synthetic_inputs(B, T, L, pinned=False)
    ls, le = label_row_bounds(L, 1, rank, world)
    y_h = (torch.rand(B, L, generator=torch.Generator().manual_seed(5)) < 0.02).float()
    x_h, len_h = onehots_h.pin_memory(), lengths_h.pin_memory()
    lab_h, yl_h = labels_h[ls:le].contiguous().pin_memory(), y_h[:, ls:le].contiguous().pin_memory()
    x_d, len_d, lab_d, y_d = x_h.to(dev), len_h.to(dev), lab_h.to(dev), yl_h.to(dev)
    loss_host = torch.zeros(1).pin_memory()
    last = {}

    def step(x, lens, lab, y):
        opt.zero_grad(set_to_none=True)
        with torch.no_grad():
            P_f = model.sequence_encoder.get_embeddings(x, lens)
        logits = pn_train.train_logits(model, P_f, lab, comm=comm, L_total=L)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, y, reduction="sum") / float(B * L)
        loss.backward()
        if comm is not None:
            pn_train.allreduce_gradients(model, comm)
        opt.step()
        last["loss"] = loss.detach()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    def step_device():
        step(x_d, len_d, lab_d, y_d)

    def step_e2e():
        step(x_h.to(dev, non_blocking=True), len_h.to(dev, non_blocking=True), lab_h.to(dev, non_blocking=True),
             yl_h.to(dev, non_blocking=True))
        loss_host.copy_(last["loss"].reshape(1), non_blocking=True)

    def whole_batch_loss():
        t = last["loss"].detach().clone().reshape(1)     # this rank's slab; the batch loss is the sum over ranks
        if world > 1:
            dist.all_reduce(t)
        return float(t)

    losses = []
    for _ in range(args.warmup):
        step_device()
        losses.append(whole_batch_loss())
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = native.launch_count()
    ms_step = timed(step_device, args.steps)
    launches = native.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    losses.append(whole_batch_loss())
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    peak_mem = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    if rank != 0:
        return
    pairs = B * L
    flops = pairs * TRAIN_FLOP_PER_PAIR + L * TRAIN_FLOP_PER_LABEL_ROW + B * TRAIN_FLOP_PER_PROTEIN
    peak, peak_src = measured_peaks()
    achieved = flops / (ms_step * 1e-3) / 1e12 / world
    passes = 3 if mode == "strict" else 1
    line = {
        "metric": "(protein,label) pairs/sec through one TRAINING step (forward + backward + Adam), batch 64 x 32K label rows",
        "value": pairs / (ms_step * 1e-3), "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32-grade (fp16 hi/lo planes, 3 tcgen05 passes)" if passes == 3 else "f16 operands, fp32 accumulate",
        "data": "synthetic",
        "config": {"workload": f"training step, batch {B} x {T} aa x {L} label rows, BCE + Adam (BASELINE.json configs[2])",
                   "mode": mode, "sequences": B, "seq_len": T, "label_rows": L,
                   "parallelism": "1 GPU" if world == 1 else f"label-sharded x{world}: BatchNorm sums and parameter "
                                                              "gradients all-reduced over NCCL",
                   "encoder": "frozen (no gradient), train-mode BatchNorm (batch statistics) as in the reference", "l2": "activations (GBs per layer) are far larger than the L2"},
        "e2e": {"value": pairs / (ms_e2e * 1e-3), "unit": "pairs/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(onehots_h.numel() * 4 + lengths_h.numel() * 8 + labels_h.numel() * 4 + y_h.numel() * 4),
                "d2h_bytes_per_step": 4 * world},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"kernel": "whole step (tensor-core GEMMs: forward, dgrad, wgrad)", "bound": "tensor",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s per GPU", "frac": achieved / peak, "traffic": None,
                     "peak_source": peak_src, "tensor_passes": passes, "executed_frac": achieved * passes / peak,
                     "algorithmic_flops_per_step": flops},
        "loss_trajectory": losses, "peak_memory_gib_rank0": peak_mem,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="strict", choices=["strict", "fast"])
    ap.add_argument("--train-mode", default=None, choices=["strict", "fast"], help="precision of --workload train")
    ap.add_argument("--sequences", type=int, default=B_TOTAL)
    ap.add_argument("--seq-len", type=int, default=T_LEN)
    ap.add_argument("--labels", type=int, default=L_ROWS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fast", action="store_true", help="skip the extra fast-mode (fp16 operand) measurement")
    ap.add_argument("--descriptions-per-label", type=int, default=1,
                    help="k consecutive label rows ensembled per label (INFERENCE_GO_DESCRIPTIONS name+label -> 2)")
    ap.add_argument("--workload", default="inference", choices=["inference", "train"],
                    help="inference = BASELINE.json configs[1] (the metric's configuration, default); "
                         "train = configs[2]: one training step, batch 64 x 32K label rows, BCE + Adam, label-sharded")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (sm_100a); there is no CPU path for the product code")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        if args.workload == "train":
            run_train(args, rank, world, local_rank)
        else:
            run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
