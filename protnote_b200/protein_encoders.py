"""B200-native mirror of the reference's `protnote/models/protein_encoders.py`.

Same class names, constructor arguments, parameter names / registration order and call signatures as the reference
(protein_encoders.py:8-153), so `state_dict()`, `load_state_dict(strict=True)`, `transfer_tf_weights_to_torch`
(protnote/utils/proteinfer.py:7-41, positional) and every caller (`bin/main.py:383-405`,
`bin/test_proteinfer.py:221-303`, `ProtNote.forward`) work unchanged.  The parameters live in ordinary torch modules;
the arithmetic of `get_embeddings` / `forward` runs in the sm_100a library (csrc/) through the C ABI.  There is no
PyTorch or CPU implementation of the forward pass in this file.
"""
from __future__ import annotations

import math

import torch

from . import native
from ._lib import ProtnoteB200Error


def _versions(tensors):
    return tuple((t.data_ptr(), t._version, tuple(t.shape), t.device.index) for t in tensors)


class MaskedConv1D(torch.nn.Conv1d):
    """Parameter holder + single-layer entry point.  Reference: protein_encoders.py:8-17
    (mask input, Conv1d(padding='same'), mask output)."""

    def forward(self, x, sequence_lengths):
        if isinstance(self.padding, str) and self.padding != "same":
            raise ProtnoteB200Error("MaskedConv1D supports padding='same' only")
        y = native.conv1d_channels_last(x, sequence_lengths, self.weight, self.bias, self.dilation[0])
        return y.permute(0, 2, 1)   # a [B, C, T] view of the channels-last result


class Residual(torch.nn.Module):
    """ResNet-v2 pre-activation bottleneck block (protein_encoders.py:23-67).  Holds parameters only; the block is
    executed as part of ProteInfer.get_embeddings (two tensor-core kernels with BatchNorm/ReLU/mask/residual fused)."""

    def __init__(self, input_channels: int, kernel_size: int, dilation: int, bottleneck_factor: float,
                 activation=torch.nn.ReLU):
        super().__init__()
        if activation is not torch.nn.ReLU:
            raise ProtnoteB200Error("the fused encoder kernels implement ReLU activations only")
        bottleneck_out_channels = int(math.floor(input_channels * bottleneck_factor))
        self.bn_activation_1 = torch.nn.Sequential(
            torch.nn.BatchNorm1d(input_channels, eps=0.001, momentum=0.01), activation())
        self.masked_conv1 = MaskedConv1D(in_channels=input_channels, out_channels=bottleneck_out_channels,
                                         padding="same", kernel_size=kernel_size, stride=1, dilation=dilation)
        self.bn_activation_2 = torch.nn.Sequential(
            torch.nn.BatchNorm1d(bottleneck_out_channels, eps=0.001, momentum=0.01), activation())
        self.masked_conv2 = MaskedConv1D(in_channels=bottleneck_out_channels, out_channels=input_channels,
                                         padding="same", kernel_size=1, stride=1, dilation=1)

    def forward(self, x, sequence_lengths):
        raise ProtnoteB200Error("Residual blocks run fused inside ProteInfer.get_embeddings; "
                                "there is no stand-alone PyTorch forward")


class ProteInfer(torch.nn.Module):
    """protein_encoders.py:70-153.  `precision`: 'strict' (fp32-grade, default) or 'fast' (fp16 operands)."""

    def __init__(self, num_labels: int, input_channels: int, output_channels: int, kernel_size: int, activation,
                 dilation_base: int, num_resnet_blocks: int, bottleneck_factor: float, precision: str = "strict"):
        super().__init__()
        self.conv1 = MaskedConv1D(in_channels=input_channels, out_channels=output_channels, padding="same",
                                  kernel_size=kernel_size, stride=1, dilation=1)
        self.resnet_blocks = torch.nn.ModuleList()
        for i in range(num_resnet_blocks):
            self.resnet_blocks.append(Residual(input_channels=output_channels, kernel_size=kernel_size,
                                               dilation=dilation_base ** i, bottleneck_factor=bottleneck_factor,
                                               activation=activation))
        self.output_layer = torch.nn.Linear(in_features=output_channels, out_features=num_labels)
        self.precision = precision
        self._dilation_base = dilation_base
        self._packed = None
        self._packed_key = None
        # training mode, sequences sharded over ranks: (process group or None for the default group, sequences of the whole
        # batch).  When set, get_embeddings() is handed THIS rank's sequences and the BatchNorm batch statistics are
        # all-reduced so that the embeddings (and the updated running statistics) are those of the unsharded batch.
        self.train_shard = None

    # ------------------------------------------------------------------ packed-weight cache
    def _pack_sources(self):
        srcs = [self.conv1.weight, self.conv1.bias]
        for blk in self.resnet_blocks:
            bn1, bn2 = blk.bn_activation_1[0], blk.bn_activation_2[0]
            srcs += [bn1.weight, bn1.bias, bn1.running_mean, bn1.running_var,
                     blk.masked_conv1.weight, blk.masked_conv1.bias,
                     bn2.weight, bn2.bias, bn2.running_mean, bn2.running_var,
                     blk.masked_conv2.weight, blk.masked_conv2.bias]
        return srcs

    def _ensure_packed(self):
        srcs = self._pack_sources()
        key = _versions(srcs) + (native.options_epoch(),)
        if self._packed is not None and self._packed_key is None:      # invalidated by a training-mode forward
            self._packed.pack(srcs)
            self._packed_key = key
        if self._packed is None or key != self._packed_key:
            bottleneck = self.resnet_blocks[0].masked_conv1.out_channels if len(self.resnet_blocks) else 1
            enc = native.PackedEncoder(self.conv1.in_channels, self.conv1.out_channels, bottleneck,
                                       self.conv1.kernel_size[0], self._dilation_base, len(self.resnet_blocks),
                                       bn_eps=self.resnet_blocks[0].bn_activation_1[0].eps if len(self.resnet_blocks) else 1e-3)
            enc.pack(srcs)
            self._packed, self._packed_key = enc, key
        return self._packed

    # ------------------------------------------------------------------ reference interface
    def get_embeddings(self, x, sequence_lengths):
        """[B, Cin, T] float + [B] lengths -> [B, C] masked mean of the residual stream (protein_encoders.py:109-118)."""
        dev = self.conv1.weight.device
        x = x.to(dev, non_blocking=True)
        sequence_lengths = sequence_lengths.to(dev, non_blocking=True)
        if self.training:
            return self._get_embeddings_train(x, sequence_lengths)
        return self._ensure_packed().forward(x, sequence_lengths, native.MODES[self.precision])

    def _get_embeddings_train(self, x, sequence_lengths):
        """.train() mode: every BatchNorm1d normalises with the statistics of this batch (over all B x T positions, padding
        included) and updates its running statistics - what the reference's frozen encoder does during training, because
        ProtNoteTrainer.train's model.train() reaches it (ProtNoteTrainer.py:844; protein_encoders.py:35-37,47-50).
        Forward only: the encoder is frozen (TRAIN_SEQUENCE_ENCODER False, base_config.yaml:71), no gradient is produced."""
        if len(self.resnet_blocks) == 0:
            return self._ensure_packed().forward(x, sequence_lengths, native.MODES[self.precision])
        if self._packed is None:
            self._ensure_packed()
        enc = self._packed
        convs = [self.conv1.weight, self.conv1.bias]
        for blk in self.resnet_blocks:
            convs += [blk.masked_conv1.weight, blk.masked_conv1.bias, blk.masked_conv2.weight, blk.masked_conv2.bias]
        # the raw pack holds conv weights only: running-statistic updates do not invalidate it
        key = _versions(convs) + (native.options_epoch(),)
        if getattr(self, "_raw_key", None) != key or getattr(enc, "packed_raw", None) is None:
            enc.pack_raw(self._pack_sources())
            self._raw_key = key
        bns, momentum = [], None
        for blk in self.resnet_blocks:
            for bn in (blk.bn_activation_1[0], blk.bn_activation_2[0]):
                if bn.momentum is None or not bn.track_running_stats:
                    raise ProtnoteB200Error("the encoder's BatchNorm layers must track running statistics with a momentum")
                momentum = bn.momentum if momentum is None else momentum
                if bn.momentum != momentum:
                    raise ProtnoteB200Error("one momentum for all encoder BatchNorm layers is assumed")
                bns += [bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var]
        group, total = self.train_shard if self.train_shard is not None else (None, None)
        out = enc.forward_train(x, sequence_lengths, bns, momentum, True, native.MODES[self.precision], group=group,
                                total_sequences=total)
        # the kernels updated the running statistics through raw pointers (torch's version counters did not move):
        # the eval-mode pack, which folds those statistics, is stale from here on
        self._packed_key = None
        for blk in self.resnet_blocks:
            for bn in (blk.bn_activation_1[0], blk.bn_activation_2[0]):
                bn.num_batches_tracked += 1
        return out

    def get_embeddings_from_tokens(self, tokens, sequence_lengths):
        """[B, T] integer residue ids (the argmax of the collator's one-hot, collators.py:123-133) -> [B, C]; the result is
        bit-identical to get_embeddings() on the one-hot, with 1 byte instead of 80 per residue crossing PCIe."""
        dev = self.conv1.weight.device
        if tokens.dtype != torch.uint8:
            # range-check BEFORE narrowing (a -1 padding id or an id >= 256 must not wrap into a valid residue); CPU
            # tensors are narrowed here so that one byte per residue crosses PCIe, CUDA tensors inside native
            if tokens.is_floating_point() or tokens.dtype == torch.bool:
                raise ValueError("token ids must be integers in [0, 255]")
            if not tokens.is_cuda:
                if tokens.numel() and (int(tokens.min()) < 0 or int(tokens.max()) > 255):
                    raise ValueError("token ids must be integers in [0, 255]")
                tokens = tokens.to(torch.uint8)
        if self.training:
            # .train() mode (batch-statistic BatchNorm, _get_embeddings_train): that kernel chain takes the one-hot layout,
            # which is expanded here ON the device - still one byte per residue over PCIe.  A pure layout expansion: ids
            # >= input_channels give an all-zero column, as in the eval-mode token kernel; padding is masked downstream.
            tokens = tokens.to(dev, non_blocking=True)
            if tokens.dtype != torch.uint8 and tokens.numel() and (int(tokens.min()) < 0 or int(tokens.max()) > 255):
                raise ValueError("token ids must be integers in [0, 255]")
            ids = torch.arange(self.conv1.in_channels, device=dev, dtype=tokens.dtype)
            onehots = (tokens[:, None, :] == ids[None, :, None]).to(torch.float32)
            return self._get_embeddings_train(onehots, sequence_lengths.to(dev, non_blocking=True))
        return self._ensure_packed().forward_tokens(tokens.to(dev, non_blocking=True),
                                                    sequence_lengths.to(dev, non_blocking=True),
                                                    native.MODES[self.precision])

    def forward(self, x, sequence_lengths):
        features = self.get_embeddings(x, sequence_lengths)
        return native.linear(features, self.output_layer.weight, self.output_layer.bias, native.MODES[self.precision])

    @classmethod
    def from_pretrained(cls, weights_path: str, num_labels: int, input_channels: int, output_channels: int,
                        kernel_size: int, activation, dilation_base: int, num_resnet_blocks: int,
                        bottleneck_factor: float):
        """protein_encoders.py:125-153.  The TF->torch weight transfer is the reference's own utility
        (protnote/utils/proteinfer.py:7-41); it zips TF variables onto state_dict() order, which this class preserves."""
        model = cls(num_labels, input_channels, output_channels, kernel_size, activation, dilation_base,
                    num_resnet_blocks, bottleneck_factor)
        from protnote.utils.proteinfer import transfer_tf_weights_to_torch  # reference package (caller's environment)
        transfer_tf_weights_to_torch(model, weights_path)
        return model
