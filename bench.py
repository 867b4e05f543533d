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
  parity    the logits of THIS run checked against the CPU oracle (oracle/, the checker): max |logit - oracle| over
            whole label rows of a few proteins, top-10 identity with the 2*tol gap guard, the columns on both sides of
            every label-shard boundary, and (N > 1) a checksum showing that every rank holds the same gathered logits.
            The model is calibrated (BatchNorm statistics matched to data, logit std ~ 2) so 1e-4 means something.
  configs   the other BASELINE.json configurations at this N (skipped with --no-configs): zero-shot EC shape
            (configs[3]), the 12-point roofline sweep (configs[4]) and the training step (configs[2], strict + fast),
            each with its value, roofline fraction and parity block
  cpu_baseline  the reference's own classes (oracle/_ref, when staged) or the oracle port timed on the host cores
  --impl reference   times that CPU implementation alone (rank 0, all host cores), same metric / unit
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
FLOP_PER_RESIDUE = 60_896_000       # encoder
FLOP_PER_LABEL_ROW = 50_331_648 + 6_291_456   # W_l + label half of output layer 1
FLOP_PER_PROTEIN = 50_798_592 + 6_291_456     # W_p + protein half of output layer 1
B_TOTAL, T_LEN, L_ROWS = 4096, 1024, 32768
EC_SHAPE = (10000, 1024, 10268, 2)   # sequences, residues, label rows, descriptions per label (BASELINE.json configs[3])
SWEEP_T, SWEEP_L, SWEEP_B = (256, 512, 1024, 2048), (1024, 8192, 32768), 256
CPU_SAMPLE = (8, 1024, 4096)        # sequences, residues, label rows of the bounded CPU sample
TOL = 1e-4                          # BASELINE.json north_star: fp32 logits within 1e-4, identical top-k


def config_tag(B, T, L, k):
    if (B, T, L, k) == (B_TOTAL, T_LEN, L_ROWS, 1):
        return "BASELINE.json configs[1]"
    if (B, T, L, k) == EC_SHAPE:
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


def synthetic_inputs(B, T, L, pinned, ragged=False):
    g = torch.Generator().manual_seed(1234)
    tokens = torch.randint(0, 20, (B, T), generator=g)
    onehots = torch.zeros(B, 20, T)
    onehots.scatter_(1, tokens[:, None, :], 1.0)
    lengths = torch.full((B,), T, dtype=torch.long)
    if ragged:
        lengths = torch.randint(max(1, T // 2), T + 1, (B,), generator=g)
        lengths[0] = T
        onehots *= (torch.arange(T)[None, :] < lengths[:, None])[:, None, :]
    labels = torch.randn(L, 1024, generator=torch.Generator().manual_seed(4321))
    if pinned:
        onehots, lengths, labels = onehots.pin_memory(), lengths.pin_memory(), labels.pin_memory()
    return onehots, lengths, labels


def calibrate_model(model, dev, world=1, logit_std=2.0):
    """Makes the random network behave like a trained one, using the PRODUCT path only (SURVEY.md section 8d asks for
    logits of std ~2: default init gives std ~1e-3 and every top-k comparison would be noise):
      1. one training-mode forward over a calibration batch with BatchNorm momentum 1 sets every running mean / variance
         (encoder, W_p, W_l, output MLP) to the statistics the layer actually sees, then they are jittered so the folded
         BatchNorm is not an exact normalisation;
      2. the output neuron is made orthogonal to the mean last-hidden activation (as oracle.synth_state_dict does for
         the golden cases), then rescaled (and its bias shifted) so the eval-mode logits of the calibration batch have
         mean 0 and std `logit_std`.
    Rank 0's result is broadcast so every rank holds bit-identical weights."""
    import torch.distributed as dist
    bns = [m for m in model.modules() if isinstance(m, torch.nn.BatchNorm1d)]
    x, lens, lab = synthetic_inputs(24, 256, 1024, pinned=False, ragged=True)
    x, lens, lab = x.to(dev), lens.to(dev), lab.to(dev)
    precision = model.precision
    model.precision = model.sequence_encoder.precision = "strict"
    saved = [m.momentum for m in bns]
    for m in bns:
        m.momentum = 1.0
    model.train()
    with torch.no_grad():
        model(sequence_onehots=x, sequence_lengths=lens, label_embeddings=lab)
    model.eval()
    g = torch.Generator().manual_seed(977)
    for m, mom in zip(bns, saved):
        m.momentum = mom
        sd = m.running_var.sqrt().cpu()
        m.running_mean.add_((0.1 * sd * torch.randn(sd.shape, generator=g)).to(dev))
        m.running_var.mul_((0.7 + 0.6 * torch.rand(sd.shape, generator=g)).to(dev)).clamp_(min=1e-4)
    with torch.no_grad():
        final = list(model.output_layer)[-1]
        # output neuron orthogonal to the mean last-hidden activation: without this the logit of a random network is a
        # huge common term cancelled by the bias (|partial sums| ~ 500 for logits of +-2), a regime no trained model is in
        # and in which fp32 rounding of the final dot product alone costs several 1e-5 - in the reference too
        _, extra = model(sequence_onehots=x, sequence_lengths=lens, label_embeddings=lab, save_embeddings=True)
        m = extra["output_layer_embeddings"].double().mean(0).to(dev)
        w = final.weight.double()[0]
        final.weight.copy_((w - (w @ m) / (m @ m) * m).float()[None, :])
        logits = model(sequence_onehots=x, sequence_lengths=lens, label_embeddings=lab)[0]
        s = logit_std / float(logits.std().clamp_min(1e-20))
        final.bias.copy_((final.bias - logits.mean()) * s)
        final.weight.mul_(s)
    model.precision = model.sequence_encoder.precision = precision
    if world > 1:
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src=0)
    with torch.no_grad():
        logits = model(sequence_onehots=x, sequence_lengths=lens, label_embeddings=lab)[0]
    model._label_cache = None
    return {"calibration_logit_std": float(logits.std()), "calibration_logit_mean": float(logits.mean()),
            "how": "product path only: train-mode forward with BatchNorm momentum 1 on 24 x 256 aa x 1024 rows, jitter, "
                   "output neuron made orthogonal to the mean hidden activation and rescaled"}


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


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own classes when oracle/_ref is staged, else the oracle port
# ------------------------------------------------------------------------------------------------------------------
def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_rate(reps: int, warmup: int):
    """pair-scores/s of the reference forward on ALL host cores.  torch.distributed.run exports OMP_NUM_THREADS=1, which
    would time one core: the thread count is set explicitly here and reported."""
    from oracle.protnote_oracle import EncoderCfg, ScorerCfg, protnote_forward, synth_inputs
    from oracle.ref_import import reference_available
    cores = host_cores()
    torch.set_num_threads(cores)
    B, T, L = CPU_SAMPLE
    ecfg, scfg = EncoderCfg(), ScorerCfg()
    model = base_config_model("strict")
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    onehots, lengths, labels = synth_inputs(B, T, L, ecfg, scfg, ragged=False, seed=1234)
    kind, impl = "port", "oracle/protnote_oracle.py (CPU restatement of the reference forward)"
    forward = lambda: protnote_forward(sd, onehots, lengths, labels, ecfg, scfg)  # noqa: E731
    if reference_available():
        try:
            from oracle.make_golden import build_reference_model
            ref = build_reference_model(ecfg, scfg, sd)

            def forward():
                with torch.no_grad():
                    return ref(sequence_onehots=onehots, sequence_lengths=lengths, label_embeddings=labels)[0]
            from oracle.ref_import import REFERENCE_ROOT
            kind, impl = "reference", ("the reference's own ProtNote / ProteInfer classes, unmodified, imported from "
                                       f"{os.path.relpath(REFERENCE_ROOT, ROOT) if REFERENCE_ROOT.startswith(ROOT) else REFERENCE_ROOT}"
                                       " (oracle/_ref is staged by oracle/build_ref.py)")
        except Exception as exc:  # noqa: BLE001 - a broken staged copy must not kill the bench: use the port, say why
            impl += f" [reference classes unavailable: {type(exc).__name__}: {exc}]"
    for _ in range(warmup):
        forward()
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        forward()
        times.append(time.perf_counter() - t0)
    mean = sum(times) / len(times)
    sample = (f"{B} seqs x {T} aa x {L} label rows per step (a bounded sample of the workload; the rate is extrapolated to "
              f"the full 4096 x 32768 step, which would take ~{B_TOTAL * L_ROWS / (B * L / mean) / 3600:.1f} h), fp32, "
              f"torch {torch.__version__} CPU ops, {impl}, {reps} timed reps, {torch.get_num_threads()} threads")
    return B * L / mean, mean * 1e3, cores, kind, sample


def run_reference(args, rank, world):
    if rank != 0:
        return
    rate, ms, cores, kind, sample = cpu_reference_rate(reps=max(1, args.steps), warmup=max(1, min(args.warmup, 1)))
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"inference {B_TOTAL} x {T_LEN} aa x {L_ROWS} label rows, fp32 in/out (BASELINE.json configs[1])",
                       "mode": "reference arithmetic, fp32, CPU", "sequences": B_TOTAL, "seq_len": T_LEN, "label_rows": L_ROWS,
                       "descriptions_per_label": 1, "parallelism": f"{cores} host cores of rank 0",
                       "sample": "each step is a bounded sample of the workload, value = sampled rate: " + sample},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "threads": torch.get_num_threads(), "kind": kind,
                             "sample": sample, "extrapolated": True},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# parity of a measured run against the CPU oracle (the checker)
# ------------------------------------------------------------------------------------------------------------------
def topk_decided_agree(ref: torch.Tensor, got: torch.Tensor, k: int, tol: float):
    """top-k indices identical at every rank the reference decides by more than 2*tol (same rule as tests/helpers.py)."""
    k = min(k, ref.shape[1])
    ri, gi = ref.topk(k, dim=1).indices, got.topk(k, dim=1).indices
    srt = ref.sort(dim=1, descending=True).values[:, :k + 1]
    clear = (srt[:, :-1] - srt[:, 1:]) > 2 * tol
    decided = torch.ones_like(ri, dtype=torch.bool)
    decided[:, :clear.shape[1]] &= clear[:, :k]
    decided[:, 1:] &= clear[:, :k - 1]
    return bool(((ri == gi) | ~decided).all()), int(decided.sum()), int((ri == gi).sum()), int(ri.numel())


def parity_block(model, logits_dev, onehots_h, lengths_h, labels_h, kdesc, world, proteins=None):
    """Rank 0: logits of the measured step vs the oracle on whole label rows of a few proteins."""
    from oracle.protnote_oracle import EncoderCfg, ScorerCfg, protnote_forward
    from protnote_b200.sharded import label_row_bounds
    t0 = time.perf_counter()
    torch.set_num_threads(host_cores())
    B, L = onehots_h.shape[0], labels_h.shape[0]
    idx = sorted(set(proteins if proteins is not None else (0, B // 2, B - 1)))
    ecfg, scfg = EncoderCfg(), ScorerCfg(inference_descriptions_per_label=kdesc)
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    sel = torch.tensor(idx)
    ref = protnote_forward(sd, onehots_h[sel].clone(), lengths_h[sel].clone(), labels_h.clone() if labels_h.is_pinned() else labels_h,
                           ecfg, scfg)
    got = logits_dev[sel.to(logits_dev.device)].float().cpu()
    err = (got - ref).abs()
    same, decided, equal, total = topk_decided_agree(ref, got, 10, TOL)
    # label columns on both sides of every shard boundary (N > 1), plus the first and last column
    cols = {0, ref.shape[1] - 1}
    for r in range(world):
        s, e = label_row_bounds(L, kdesc, r, world)
        for c in (s // kdesc - 1, s // kdesc, e // kdesc - 1, e // kdesc):
            if 0 <= c < ref.shape[1]:
                cols.add(c)
    cols = sorted(cols)
    return {"checker": "oracle/protnote_oracle.py (fp32 CPU restatement of the reference forward, pinned against the "
                       "reference classes and tests/golden)", "proteins": idx, "label_columns": int(ref.shape[1]),
            "pairs_checked": int(ref.numel()), "max_abs_err": float(err.max()), "mean_abs_err": float(err.mean()),
            "tol": TOL, "within_tol": bool(err.max() <= TOL), "logit_std": float(ref.std()), "logit_absmax": float(ref.abs().max()),
            "top10_identical_where_decided": same, "top10_ranks_decided": decided, "top10_ranks_equal": equal,
            "top10_ranks_total": total, "top10_gap_guard": 2 * TOL,
            "shard_boundary_columns": cols, "shard_boundary_max_abs_err": float(err[:, cols].max()),
            "seconds": time.perf_counter() - t0}


def ranks_agree(logits_dev, world):
    """Every rank holds the same gathered logits: position-weighted fp64 checksums compared across ranks."""
    import torch.distributed as dist
    v = logits_dev.double().flatten()
    idx = torch.arange(v.numel(), dtype=torch.int64, device=v.device)
    w = ((idx * 2654435761) & 0xFFFFFFFF).double() / 4294967296.0 + 0.5
    digest = torch.stack([v.sum(), v.abs().sum(), (v * w).sum()])
    if world == 1:
        return {"ranks": 1, "identical": True, "checksum": [float(t) for t in digest]}
    allv = [torch.empty_like(digest) for _ in range(world)]
    dist.all_gather(allv, digest)
    allv = torch.stack(allv)
    return {"ranks": world, "identical": bool((allv == allv[0:1]).all()),
            "max_checksum_spread": float((allv - allv[0:1]).abs().max()), "checksum": [float(t) for t in digest]}


# ------------------------------------------------------------------------------------------------------------------
# one inference workload, device-resident timing (used by the headline and by the EC / sweep configurations)
# ------------------------------------------------------------------------------------------------------------------
class Inference:
    def __init__(self, model, dev, rank, world, B, T, L, kdesc, pinned=True):
        from protnote_b200.sharded import label_row_bounds, shard_bounds
        self.model, self.dev, self.rank, self.world = model, dev, rank, world
        self.B, self.T, self.L, self.k = B, T, L, kdesc
        if model.inference_descriptions_per_label != kdesc:
            model.inference_descriptions_per_label = kdesc        # part of the pack key: the scorer re-packs
        self.onehots_h, self.lengths_h, self.labels_h = synthetic_inputs(B, T, L, pinned=pinned and world == 1)
        ps, pe = shard_bounds(B, rank, world)
        self.ls, self.le = label_row_bounds(L, kdesc, rank, world)
        self.x_h, self.len_h, self.lab_h = self.onehots_h[ps:pe], self.lengths_h[ps:pe], self.labels_h[self.ls:self.le]
        if world > 1 and pinned:
            self.x_h, self.len_h, self.lab_h = (t.contiguous().pin_memory() for t in (self.x_h, self.len_h, self.lab_h))
        self.x_d, self.len_d, self.lab_d = self.x_h.to(dev), self.len_h.to(dev), self.lab_h.to(dev)
        self.out = None

    def forward(self, x, lens, lab):
        from protnote_b200.sharded import native_sharded_forward
        self.model._label_cache = None     # W_l(label_embeddings) is recomputed every step, like the reference does
        with torch.no_grad():
            if self.world == 1:
                return self.model(sequence_onehots=x, sequence_lengths=lens, label_embeddings=lab)[0]
            return native_sharded_forward(self.model, x, lens, lab, inputs_are_local=True, total_sequences=self.B,
                                          total_label_rows=self.L)

    def step_device(self):
        self.out = self.forward(self.x_d, self.len_d, self.lab_d)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    def flops(self):
        return (self.B * self.L * FLOP_PER_PAIR + self.B * self.T * FLOP_PER_RESIDUE + self.L * FLOP_PER_LABEL_ROW
                + self.B * FLOP_PER_PROTEIN)


def release_memory():
    from protnote_b200 import native
    native.release_scratch()
    torch.cuda.empty_cache()


def run_small_config(model, dev, rank, world, B, T, L, kdesc, steps, warmup, with_parity):
    """value / whole-step roofline fraction (+ parity) of one non-headline inference configuration."""
    from protnote_b200 import native
    inf = Inference(model, dev, rank, world, B, T, L, kdesc, pinned=False)
    for _ in range(warmup):
        inf.step_device()
    native.gemm_timing(True)
    ms = inf.timed(inf.step_device, steps)
    gemm_ms, gemm_n, gemm_flops = native.gemm_timing_read()
    native.gemm_timing(False)
    agree = ranks_agree(inf.out, world)
    res = None
    if rank == 0:
        peak, _ = measured_peaks()
        tf = inf.flops() / (ms * 1e-3) / 1e12 / world
        res = {"sequences": B, "seq_len": T, "label_rows": L, "descriptions_per_label": kdesc, "value": B * L / (ms * 1e-3),
               "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": warmup,
               "algorithmic_tflops_per_gpu": tf, "frac_of_peak": tf / peak,
               "scorer_gemm_tflops": gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None,
               "outputs_finite": bool(torch.isfinite(inf.out).all().item()), "ranks_agree": agree["identical"]}
        if with_parity:
            res["parity"] = parity_block(model, inf.out, inf.onehots_h, inf.lengths_h, inf.labels_h, kdesc, world)
    del inf
    release_memory()
    return res


# ------------------------------------------------------------------------------------------------------------------
# headline
# ------------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    from protnote_b200 import native

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    B, T, L = args.sequences, args.seq_len, args.labels
    kdesc = args.descriptions_per_label
    model = base_config_model(args.mode, kdesc).to(dev)
    calib = calibrate_model(model, dev, world)
    inf = Inference(model, dev, rank, world, B, T, L, kdesc, pinned=True)
    ls, le = inf.ls, inf.le
    logits_host = torch.empty(B, (le - ls) // kdesc, dtype=torch.float32).pin_memory()

    def step_e2e():
        x = inf.x_h.to(dev, non_blocking=True)
        lens = inf.len_h.to(dev, non_blocking=True)
        lab = inf.lab_h.to(dev, non_blocking=True)
        inf.out = inf.forward(x, lens, lab)
        logits_host.copy_(inf.out[:, ls // kdesc:le // kdesc] if world > 1 else inf.out, non_blocking=True)

    for _ in range(args.warmup):
        inf.step_device()
    inf.barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = native.launch_count()
    native.gemm_timing(True)
    ms_step = inf.timed(inf.step_device, args.steps)
    gemm_ms, gemm_launches, gemm_flops = native.gemm_timing_read()
    native.gemm_timing(False)
    launches = native.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    out = inf.out
    finite = bool(torch.isfinite(out).all().item())
    agree = ranks_agree(out, world)
    parity = None
    if not args.no_parity and rank == 0:
        parity = parity_block(model, out, inf.onehots_h, inf.lengths_h, inf.labels_h, kdesc, world)
        parity["ranks_agree"] = agree
    inf.barrier()

    step_e2e()                              # one untimed e2e step (pinned-buffer / allocator warm-up)
    ms_e2e = inf.timed(step_e2e, args.steps)
    e2e_max_diff = float((logits_host.to(dev) - (out[:, ls // kdesc:le // kdesc] if world > 1 else out)).abs().max())

    # per-stage device times of one step on this rank's shard (outside the timed regions; explains `value`)
    breakdown = {}
    with torch.no_grad():
        scorer = model._ensure_packed()
        mode = native.MODES[args.mode]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        torch.cuda.synchronize()
        ev[0].record()
        P_f = model.sequence_encoder.get_embeddings(inf.x_d, inf.len_d)
        ev[1].record()
        _, a = scorer.project_sequences(P_f, mode)
        ev[2].record()
        _, c = scorer.project_labels(inf.lab_d, mode)
        ev[3].record()
        scorer.score(a, c, mode=mode)
        ev[4].record()
        torch.cuda.synchronize()
        for i, name in enumerate(("encoder", "W_p+layer1_p", "W_l+layer1_l", "pair_scorer")):
            breakdown[name + "_ms"] = ev[i].elapsed_time(ev[i + 1])
        breakdown["encoder_residues_per_s"] = inf.x_d.shape[0] * T / (breakdown["encoder_ms"] * 1e-3)
        breakdown["encoder_algorithmic_tflops"] = inf.x_d.shape[0] * T * FLOP_PER_RESIDUE / (breakdown["encoder_ms"] * 1e-3) / 1e12
        breakdown["scorer_pairs_per_s"] = a.shape[0] * c.shape[0] / (breakdown["pair_scorer_ms"] * 1e-3)
        del P_f, a, c

    # the same step with one tensor-core pass per product (fp16 operands, fp32 accumulate = the arithmetic the reference
    # itself uses on a GPU under torch.autocast): how fast the kernels are when fp32-grade logits are not required, and
    # how far those logits are from the strict ones
    fast = None
    if args.mode == "strict" and world == 1 and not args.no_fast:
        strict_out = out
        model.precision = model.sequence_encoder.precision = "fast"
        inf.step_device()
        native.gemm_timing(True)
        ms_fast = inf.timed(inf.step_device, 1)
        f_ms, f_n, f_flops = native.gemm_timing_read()
        native.gemm_timing(False)
        model.precision = model.sequence_encoder.precision = "strict"
        diff = (inf.out - strict_out).abs()
        k10 = min(10, out.shape[1])
        same = (inf.out.topk(k10, dim=1).indices == strict_out.topk(k10, dim=1).indices).all(dim=1).float().mean()
        fast = {"value": B * L / (ms_fast * 1e-3), "unit": UNIT, "ms_per_step": ms_fast,
                "scorer_gemm_tflops": f_flops / (f_ms * 1e-3) / 1e12 if f_ms > 0 else None,
                "max_abs_logit_diff_vs_strict": float(diff.max()), "mean_abs_logit_diff_vs_strict": float(diff.mean()),
                "logit_std": float(strict_out.std()), "top10_identical_fraction": float(same),
                "max_abs_logit_diff_over_logit_std": float(diff.max() / strict_out.std().clamp_min(1e-30)),
                "note": "NOT the headline: fp16-operand arithmetic (what the reference uses on a GPU under torch.autocast) "
                        "is a few percent of the logit std away from the fp32 result, far outside the 1e-4 bar"}
        inf.out = strict_out
        del diff
    h2d = inf.onehots_h.numel() * 4 + inf.lengths_h.numel() * 8 + inf.labels_h.numel() * 4
    d2h = B * (L // kdesc) * 4
    headline_flops = inf.flops()
    del inf, out, logits_host
    release_memory()

    # ---- the other BASELINE.json configurations at this N
    configs = None
    if not args.no_configs and args.mode == "strict" and ((B, T, L, kdesc) == (B_TOTAL, T_LEN, L_ROWS, 1) or args.force_configs):
        configs = run_other_configs(args, model, dev, rank, world)
    if rank != 0:
        return
    pairs = B * L
    peak, peak_src = measured_peaks()
    passes = 3 if args.mode == "strict" else 1
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
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
                   "label_projection": "recomputed every step (cache cleared)",
                   "weights": "random init of the published architecture, calibrated to logit std ~2", "calibration": calib},
        "e2e": {"value": pairs / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "max_abs_diff_vs_device_resident_logits": e2e_max_diff},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "parity": parity,
        "roofline": {"kernel": "pn::gemm_kernel<32,3> (pair-scorer output-MLP layers 2 and 3)" if passes == 3
                     else "pn::gemm_kernel<64,1>", "bound": "tensor", "achieved": achieved, "peak": peak,
                     "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                     # NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum per launch of the committed
                     # `ncu --set full` capture of round 2 (2^19-row launches: layer 2 reads 16.78 GB + writes 4.63 GB =
                     # 40.8 KB/row, the dot-epilogue layer 14.35 + 0.69 GB = 28.7 KB/row -> 34.8 KB per row per launch),
                     # scaled by the rows of the timed launches; algorithmic: 12 KB/row of A planes read, +12 KB/row written
                     # by the storing layer = 18 KB/row.  The excess is the 38 MB of weights not staying L2-resident.
                     "traffic": (34.8e3 * gemm_flops / (2.0 * 3072 * 3072) / gemm_launches) if gemm_launches else None,
                     "traffic_source": "profiles/r02_ncu_full_after.txt (constant from that ncu capture, scaled by rows; "
                                       "not re-measured per run)",
                     "traffic_unit": "bytes per launch",
                     "algorithmic_bytes_per_launch": (18.0e3 * gemm_flops / (2.0 * 3072 * 3072) / gemm_launches)
                     if gemm_launches else None,
                     "peak_source": peak_src, "launches": int(gemm_launches),
                     "avg_launch_ms": gemm_ms / gemm_launches if gemm_launches else None,
                     "share_of_step": gemm_ms / (ms_step * args.steps) if ms_step > 0 else None,
                     "tensor_passes": passes, "executed_tflops": achieved * passes,
                     "executed_frac": achieved * passes / peak if peak else None,
                     "whole_step_algorithmic_tflops_per_gpu": headline_flops / (ms_step * 1e-3) / 1e12 / world,
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
    if configs is not None:
        line["configs"] = configs
    if world == 1 and not args.no_cpu_baseline:
        rate, ms, cores, kind, sample = cpu_reference_rate(reps=2, warmup=1)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "threads": torch.get_num_threads(), "kind": kind,
                                "sample": sample, "ms_per_sample": ms, "extrapolated": True}
    print(json.dumps(line), flush=True)


def run_other_configs(args, model, dev, rank, world):
    """BASELINE.json configs[3] (zero-shot EC shape), configs[4] (roofline sweep) and configs[2] (training step) at this N.
    Every rank runs them (the sharded paths are collective); rank 0 keeps the results."""
    out = {"note": "measured after the headline in the same process, device-resident inputs, CUDA events, max over ranks; "
                   "fewer steps than the headline (stated per entry)"}
    t0 = time.perf_counter()
    out["memory_allocated_gib_after_headline"] = torch.cuda.memory_allocated(dev) / 2 ** 30
    try:
        B, T, L, k = EC_SHAPE
        out["ec"] = run_small_config(model, dev, rank, world, B, T, L, k, steps=2, warmup=1, with_parity=True)
        if rank == 0:
            out["ec"]["workload"] = config_tag(B, T, L, k)
    except Exception as exc:  # noqa: BLE001 - an auxiliary configuration must not lose the headline line
        out["ec"] = {"error": f"{type(exc).__name__}: {exc}"}
    release_memory()
    out["memory_allocated_gib_after_ec"] = torch.cuda.memory_allocated(dev) / 2 ** 30
    sweep = []
    try:
        for T in SWEEP_T:
            for L in SWEEP_L:
                r = run_small_config(model, dev, rank, world, SWEEP_B, T, L, 1, steps=2, warmup=1,
                                     with_parity=(T, L) == (SWEEP_T[0], SWEEP_L[0]))
                if rank == 0:
                    keep = ("seq_len", "label_rows", "value", "ms_per_step", "algorithmic_tflops_per_gpu", "frac_of_peak",
                            "scorer_gemm_tflops", "outputs_finite", "ranks_agree", "parity")
                    sweep.append({k2: r[k2] for k2 in keep if k2 in r})
        out["sweep"] = {"sequences": SWEEP_B, "unit": UNIT, "points": sweep,
                        "workload": "BASELINE.json configs[4]: seq-len {256,512,1024,2048} x label rows {1K,8K,32K}"}
    except Exception as exc:  # noqa: BLE001
        out["sweep"] = {"error": f"{type(exc).__name__}: {exc}", "points": sweep}
    model.inference_descriptions_per_label = 1
    # the inference model's packs, label cache and every cached workspace go before the training step needs ~100 GB
    model._packed = model._packed_key = model._label_cache = None
    model.sequence_encoder._packed = model.sequence_encoder._packed_key = None
    del model
    release_memory()
    out["memory_allocated_gib_before_train"] = torch.cuda.memory_allocated(dev) / 2 ** 30
    try:
        out["train"] = train_configs(args, dev, rank, world)
    except Exception as exc:  # noqa: BLE001
        out["train"] = {"error": f"{type(exc).__name__}: {exc}"}
    out["seconds"] = time.perf_counter() - t0
    return out if rank == 0 else None


# ------------------------------------------------------------------------------------------------------------------
# training step (BASELINE.json configs[2])
# ------------------------------------------------------------------------------------------------------------------
TRAIN_B = 64
# algorithmic FLOPs of one training step per (protein, label-row) pair: forward 2 GEMMs + dot, backward dgrad + wgrad of
# the same two layers (layer 1 is factorised in both directions) -> 3 x the forward GEMM work
TRAIN_FLOP_PER_PAIR = 3 * 2 * (2 * 3072 * 3072) + 2 * 3072
TRAIN_FLOP_PER_LABEL_ROW = 3 * 50_331_648 + 3 * 2 * 1024 * 3072     # W_l forward + backward, label half of layer 1
TRAIN_FLOP_PER_PROTEIN = 3 * 50_798_592 + 3 * 2 * 1024 * 3072 + 1024 * 60_896_000   # W_p, protein half, frozen encoder
# activation memory of a strict step at batch 64: measured 101.6 GiB at 16 384 rows per rank (2 GPUs) incl. ~12 GiB fixed
TRAIN_STRICT_BYTES_PER_LABEL_ROW = 180 * 2 ** 30 / 32768


class TrainStep:
    """One training step = frozen encoder forward (proteins sharded over ranks, BatchNorm sums all-reduced) -> W_p / W_l /
    output MLP forward with BATCH statistics -> fused loss + gradient seed -> backward -> gradient all-reduce of the
    label-sharded ranks -> Adam."""

    def __init__(self, dev, rank, world, mode, B, T, L, loss="bce"):
        from protnote_b200 import train as pn_train
        from protnote_b200.sharded import label_row_bounds
        self.pn_train, self.dev, self.rank, self.world, self.B, self.T, self.L = pn_train, dev, rank, world, B, T, L
        self.mode, self.loss_name = mode, loss
        # model.train() as ProtNoteTrainer.train does it: the frozen encoder's BatchNorm layers use batch statistics too
        self.model = base_config_model(mode).to(dev).train()
        self.params = pn_train.trainable_parameters(self.model)
        for p in self.model.sequence_encoder.parameters():
            p.requires_grad_(False)
        self.opt = torch.optim.Adam(self.params, lr=3e-4, fused=True)
        self.comm = pn_train.Comm() if world > 1 else None
        onehots_h, lengths_h, labels_h = synthetic_inputs(B, T, L, pinned=False)
        ls, le = label_row_bounds(L, 1, rank, world)
        y_h = (torch.rand(B, L, generator=torch.Generator().manual_seed(5)) < 0.02).float()
        from protnote_b200.sharded import shard_bounds
        ps, pe = shard_bounds(B, rank, world)      # the frozen encoder runs on this rank's proteins only
        self.x_h, self.len_h = onehots_h[ps:pe].contiguous().pin_memory(), lengths_h[ps:pe].contiguous().pin_memory()
        if world > 1:
            self.model.sequence_encoder.train_shard = (None, B)
        self.lab_h, self.yl_h = labels_h[ls:le].contiguous().pin_memory(), y_h[:, ls:le].contiguous().pin_memory()
        self.h2d = onehots_h.numel() * 4 + lengths_h.numel() * 8 + labels_h.numel() * 4 + y_h.numel() * 4
        self.x_d, self.len_d, self.lab_d, self.y_d = (t.to(dev) for t in (self.x_h, self.len_h, self.lab_h, self.yl_h))
        self.loss_host = torch.zeros(1).pin_memory()
        self.last_loss = None

    def step(self, x, lens, lab, y):
        pn_train = self.pn_train
        self.opt.zero_grad(set_to_none=True)
        with torch.no_grad():
            P_f = self.model.sequence_encoder.get_embeddings(x, lens)
            if self.world > 1:      # [B / W, C] -> [B, C] on every rank (18 KB per protein)
                from protnote_b200.sharded import all_gather_rows
                P_f = all_gather_rows(P_f, self.B)
        # BCE evaluated inside the last forward kernel (loss + gradient seed: the [B, L] logits never round-trip through
        # autograd); the parameter gradients are all-reduced inside the backward, overlapped with its GEMMs
        loss, _ = pn_train.train_loss(self.model, P_f, lab, y, loss=self.loss_name, comm=self.comm, L_total=self.L,
                                      reduce_gradients=True)
        loss.backward()
        self.opt.step()
        self.last_loss = loss.detach()

    def step_device(self):
        self.step(self.x_d, self.len_d, self.lab_d, self.y_d)

    def step_e2e(self):
        dev = self.dev
        self.step(self.x_h.to(dev, non_blocking=True), self.len_h.to(dev, non_blocking=True),
                  self.lab_h.to(dev, non_blocking=True), self.yl_h.to(dev, non_blocking=True))
        self.loss_host.copy_(self.last_loss.reshape(1), non_blocking=True)

    def whole_batch_loss(self):
        t = self.last_loss.detach().clone().reshape(1)     # this rank's slab; the batch loss is the sum over ranks
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t)
        return float(t)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    def flops(self):
        return self.B * self.L * TRAIN_FLOP_PER_PAIR + self.L * TRAIN_FLOP_PER_LABEL_ROW + self.B * TRAIN_FLOP_PER_PROTEIN


def train_measure(dev, rank, world, mode, B, T, L, steps, warmup, e2e=True):
    from protnote_b200 import native
    torch.cuda.reset_peak_memory_stats(dev)
    ts = TrainStep(dev, rank, world, mode, B, T, L)
    losses = []
    for _ in range(warmup):
        ts.step_device()
        losses.append(ts.whole_batch_loss())
    sampler = ClockSampler(dev.index) if rank == 0 else None
    launches0 = native.launch_count()
    ms_step = ts.timed(ts.step_device, steps)
    launches = native.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    losses.append(ts.whole_batch_loss())
    ms_e2e = None
    if e2e:
        ts.step_e2e()
        ms_e2e = ts.timed(ts.step_e2e, steps)
    peak_mem = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    flops, h2d = ts.flops(), ts.h2d
    del ts
    release_memory()
    return dict(ms_step=ms_step, ms_e2e=ms_e2e, launches=launches, clocks=clocks, losses=losses, peak_mem=peak_mem,
                flops=flops, h2d=h2d)


def train_parity_small(dev, rank, world):
    """The label-sharded training step on THIS hardware and THIS process group (NCCL) against the single-process CPU
    training oracle (oracle/train_oracle.py, pinned against the reference class in train mode): logits, loss and every
    parameter gradient of the golden case 'train_tiny_wide' (3 proteins x 130 label rows, strict arithmetic)."""
    from oracle.make_golden_train import train_inputs
    from oracle.train_oracle import train_step_oracle
    from protnote_b200 import train as pn_train
    from protnote_b200.sharded import all_gather_columns, label_row_bounds
    from tests.helpers import build_b200_model
    ecfg, scfg, sd, P_f, L_f, y = train_inputs("train_tiny_wide")
    B, L = y.shape
    model = build_b200_model(ecfg, scfg, sd, device=dev).train()
    comm = pn_train.Comm() if world > 1 else None
    ls, le = label_row_bounds(L, 1, rank, world)
    # the measured path: loss fused into the last forward kernel, gradients all-reduced inside the backward
    loss, logits = pn_train.train_loss(model, P_f.to(dev), L_f[ls:le].to(dev).contiguous(), y[:, ls:le].to(dev).contiguous(),
                                       loss="bce", comm=comm, L_total=L, reduce_gradients=True)
    loss.backward()
    full = all_gather_columns(logits.detach(), L) if world > 1 else logits.detach()
    tot = loss.detach().clone().reshape(1)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(tot)
    # the frozen encoder in train mode with the SEQUENCES sharded over the ranks (BatchNorm sums all-reduced)
    from oracle.train_oracle import proteinfer_embeddings_train
    from protnote_b200.sharded import all_gather_rows, shard_bounds
    from tests.helpers import load_case
    e_ecfg, e_scfg, e_sd, onehots, lengths, _, _ = load_case("tiny_concat")
    enc = build_b200_model(e_ecfg, e_scfg, e_sd, device=dev).train().sequence_encoder
    ps, pe = shard_bounds(onehots.shape[0], rank, world)
    if world > 1:
        enc.train_shard = (None, onehots.shape[0])
    with torch.no_grad():
        emb = enc.get_embeddings(onehots[ps:pe].to(dev), lengths[ps:pe].to(dev))
        emb = all_gather_rows(emb, onehots.shape[0]) if world > 1 else emb
    if rank != 0:
        return None
    torch.set_num_threads(host_cores())
    o_emb, o_stats = proteinfer_embeddings_train(e_sd, onehots, lengths, e_ecfg)
    enc_err = float((emb.cpu().double() - o_emb).abs().max())
    bufs = dict(enc.named_buffers())
    stat_err = max(float((bufs[k[len("sequence_encoder."):]].cpu().double() - v).abs().max()) for k, v in o_stats.items())
    o_logits, o_loss, o_grads, _ = train_step_oracle(sd, P_f, L_f, y, scfg)
    named = dict(model.named_parameters())
    worst_rel, worst_max = 0.0, 0.0
    for k, g in o_grads.items():
        d = named[k].grad.cpu().double() - g
        worst_rel = max(worst_rel, float(d.norm() / g.norm().clamp_min(1e-30)))
        worst_max = max(worst_max, float(d.abs().max() / g.abs().max().clamp_min(1e-30)))
    return {"checker": "oracle/train_oracle.py (fp64 autograd restatement, pinned against the reference ProtNote class in "
                       "train mode), golden case train_tiny_wide (3 x 130 label rows), strict arithmetic",
            "ranks": world, "max_abs_logit_err": float((full.cpu().double() - o_logits).abs().max()),
            "loss": float(tot), "oracle_loss": float(o_loss), "abs_loss_err": abs(float(tot) - float(o_loss)),
            "gradients_checked": len(o_grads), "worst_gradient_rel_l2_err": worst_rel,
            "worst_gradient_max_err_over_max_entry": worst_max,
            "encoder_train_mode": {"what": "frozen encoder, BatchNorm batch statistics, sequences sharded over the ranks "
                                           "(golden case tiny_concat) vs oracle.train_oracle.proteinfer_embeddings_train (fp64)",
                                   "max_abs_embedding_err": enc_err, "max_abs_running_stat_err": stat_err},
            "within_tol": bool((full.cpu().double() - o_logits).abs().max() <= TOL and worst_max <= 1e-3
                               and enc_err <= 1e-5 and stat_err <= 1e-5)}


def train_line(r, mode, world, B, T, L, steps, warmup):
    peak, peak_src = measured_peaks()
    pairs = B * L
    achieved = r["flops"] / (r["ms_step"] * 1e-3) / 1e12 / world
    passes = 3 if mode == "strict" else 1
    return {"value": pairs / (r["ms_step"] * 1e-3), "unit": "pairs/s", "ms_per_step": r["ms_step"], "steps": steps, "warmup": warmup,
            "mode": mode, "e2e_ms_per_step": r["ms_e2e"], "algorithmic_tflops_per_gpu": achieved, "frac_of_peak": achieved / peak,
            "executed_frac": achieved * passes / peak, "tensor_passes": passes, "gpu_launches": int(r["launches"]),
            "loss_trajectory": r["losses"], "peak_memory_gib_rank0": r["peak_mem"]}


def train_configs(args, dev, rank, world):
    """configs[2] at this N: fast and (where the activations fit) strict, plus the small on-hardware parity step."""
    T, L, B = T_LEN, L_ROWS, TRAIN_B
    out = {"workload": f"training step, batch {B} x {T} aa x {L} label rows, BCE + Adam (BASELINE.json configs[2]), "
                       f"label-sharded x{world}"}
    par = train_parity_small(dev, rank, world)
    if rank == 0:
        out["parity"] = par
    release_memory()
    for mode in ("fast", "strict"):
        if mode == "strict":
            need = TRAIN_STRICT_BYTES_PER_LABEL_ROW * (L / world) + 12 * 2 ** 30
            free = torch.cuda.mem_get_info(dev)[0]
            if need > 0.85 * free:
                out[mode] = {"skipped": f"strict activations of {L // world} label rows per rank need ~{need / 2 ** 30:.0f} GiB "
                                        f"(> 85% of the {free / 2 ** 30:.0f} GiB free); runs label-sharded on 2 or more GPUs"}
                continue
        try:
            r = train_measure(dev, rank, world, mode, B, T, L, steps=3, warmup=2, e2e=False)
            if rank == 0:
                out[mode] = train_line(r, mode, world, B, T, L, 3, 2)
        except Exception as exc:  # noqa: BLE001 - keep what was measured
            out[mode] = {"error": f"{type(exc).__name__}: {str(exc)[:300]}",
                         "memory_allocated_gib": torch.cuda.memory_allocated(dev) / 2 ** 30}
            release_memory()
    return out if rank == 0 else None


def run_train(args, rank, world, local_rank):
    """BASELINE.json configs[2] as the main line (--workload train)."""
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    B = TRAIN_B if args.sequences == B_TOTAL else args.sequences
    T, L = args.seq_len, args.labels
    mode = args.mode if args.train_mode is None else args.train_mode
    par = None if args.no_parity else train_parity_small(dev, rank, world)
    r = train_measure(dev, rank, world, mode, B, T, L, args.steps, args.warmup, e2e=True)
    if rank != 0:
        return
    pairs = B * L
    peak, peak_src = measured_peaks()
    achieved = r["flops"] / (r["ms_step"] * 1e-3) / 1e12 / world
    passes = 3 if mode == "strict" else 1
    line = {
        "metric": "(protein,label) pairs/sec through one TRAINING step (forward + backward + Adam), batch 64 x 32K label rows",
        "value": pairs / (r["ms_step"] * 1e-3), "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32-grade (fp16 hi/lo planes, 3 tcgen05 passes)" if passes == 3 else "f16 operands, fp32 accumulate",
        "data": "synthetic",
        "config": {"workload": f"training step, batch {B} x {T} aa x {L} label rows, BCE + Adam (BASELINE.json configs[2])",
                   "mode": mode, "sequences": B, "seq_len": T, "label_rows": L,
                   "parallelism": "1 GPU" if world == 1 else f"label-sharded x{world}: BatchNorm sums and parameter "
                                                              "gradients all-reduced over NCCL",
                   "encoder": "frozen (no gradient), train-mode BatchNorm (batch statistics) as in the reference", "l2": "activations (GBs per layer) are far larger than the L2"},
        "e2e": {"value": pairs / (r["ms_e2e"] * 1e-3), "unit": "pairs/s", "ms_per_step": r["ms_e2e"],
                "h2d_bytes_per_step": int(r["h2d"]), "d2h_bytes_per_step": 4 * world},
        "gpu_launches": int(r["launches"]), "clocks": r["clocks"], "parity": par,
        "roofline": {"kernel": "whole step (tensor-core GEMMs: forward, dgrad, wgrad)", "bound": "tensor",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s per GPU", "frac": achieved / peak, "traffic": None,
                     "peak_source": peak_src, "tensor_passes": passes, "executed_frac": achieved * passes / peak,
                     "algorithmic_flops_per_step": r["flops"]},
        "loss_trajectory": r["losses"], "peak_memory_gib_rank0": r["peak_mem"],
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
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run check of the logits against the CPU oracle")
    ap.add_argument("--no-configs", action="store_true",
                    help="skip the other BASELINE.json configurations (EC shape, sweep, training step) after the headline")
    ap.add_argument("--force-configs", action="store_true", help="run the `configs` block after a non-headline shape too (debugging)")
    ap.add_argument("--descriptions-per-label", type=int, default=1,
                    help="k consecutive label rows ensembled per label (INFERENCE_GO_DESCRIPTIONS name+label -> 2)")
    ap.add_argument("--workload", default="inference", choices=["inference", "train", "ec"],
                    help="inference = BASELINE.json configs[1] (the metric's configuration, default; followed by a compact "
                         "`configs` block with configs[2], [3], [4] at this N); ec = configs[3] as the main line; "
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
    if args.workload == "ec":
        args.sequences, args.seq_len, args.labels, args.descriptions_per_label = EC_SHAPE
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        # a rank-order bug must fail in minutes, not sit in a collective until the default 10-minute watchdog fires
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=300))
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
