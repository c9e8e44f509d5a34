"""Data-parallel pieces of the training step (BASELINE config 5).

The reference wraps the models in `nn.DataParallel` (train/adversarial_wasserstein_gp.py:192-196): one process
scatters the batch over the GPUs, gathers the outputs and reduces the gradients on GPU 0 every step.  Here the
layout is one process per GPU (torchrun) with the batch sharded by rank and ONE collective per step: the
gradients of the step are packed into a few flat buckets and all-reduced (sum) over NCCL -- 24,073,584 elements
for a generator step, 119,233 for a discriminator step -- then divided by the world size, which reproduces the
gradient of the mean loss over the global batch.  Text2Mel and the discriminator use LayerNorm only, so there
are no batch statistics to synchronise.

    for step ...:
        loss = generator_loss(m1(mel, ids, spk), ...)      # spoofsv_b200 kernels, autograd graph
        loss.backward()
        allreduce_gradients(m1.parameters())                # NCCL over NVLink, bucketed
        optimizer.step()
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist


def plan_buckets(numels: Sequence[int], bucket_elems: int) -> List[List[int]]:
    """Greedy, order-preserving packing of tensors (by index) into buckets of at most bucket_elems elements;
    a tensor larger than the limit gets a bucket of its own."""
    if bucket_elems < 1:
        raise ValueError("bucket_elems must be positive")
    buckets, cur, used = [], [], 0
    for i, n in enumerate(numels):
        if cur and used + n > bucket_elems:
            buckets.append(cur)
            cur, used = [], 0
        cur.append(i)
        used += n
    if cur:
        buckets.append(cur)
    return buckets


def allreduce_gradients(params: Iterable[torch.nn.Parameter], bucket_mb: float = 32.0, average: bool = True,
                        group=None) -> int:
    """Sum the `.grad` of every parameter over the ranks of the process group (in place), bucketed into flat
    buffers of about bucket_mb megabytes; divides by the world size when `average`.  Parameters without a
    gradient on this rank contribute zeros (every rank must pass the same parameter list).  Returns the number of
    elements reduced.  A single-process run (no process group) is a no-op."""
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    world = dist.get_world_size(group)
    plist = [p for p in params if p.requires_grad]
    if world == 1 or not plist:
        return 0
    for p in plist:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    total = 0
    by_kind = {}
    for p in plist:
        by_kind.setdefault((p.grad.device, p.grad.dtype), []).append(p)
    for (device, dtype), ps in by_kind.items():
        elem = torch.empty((), dtype=dtype).element_size()
        limit = max(1, int(bucket_mb * (1 << 20) / elem))
        handles = []
        for idxs in plan_buckets([p.grad.numel() for p in ps], limit):
            grads = [ps[i].grad for i in idxs]
            flat = torch.cat([g.reshape(-1) for g in grads])
            handles.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True), flat, grads))
            total += flat.numel()
        for work, flat, grads in handles:           # buckets are in flight together; unpack in order
            work.wait()
            if average:
                flat.div_(world)
            off = 0
            for g in grads:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g))
                off += n
    return total


class OverlappedGradReducer:
    """Gradient allreduce overlapped with the backward pass (replaces the reduce-add onto GPU 0 that `nn.DataParallel`
    performs after backward, train/adversarial_wasserstein_gp.py:183-196).

    The parameters are packed, in REVERSE registration order (the order their gradients become final during backward:
    AudioDec, AudioEnc, then TextEnc, whose 68 MB finish last), into buckets of about `bucket_mb`; a
    post-accumulate-grad hook counts the finished gradients of a bucket and launches its NCCL allreduce the moment the
    last one lands, so the collective of layer k runs under the backward kernels of the layers before it.  `finish()`
    after `backward()` launches whatever never became ready (parameters without a gradient contribute zeros), waits,
    averages and scatters the sums back into `.grad`.  One backward pass per step (the generator iteration); the
    discriminator iteration calls backward twice and keeps `allreduce_gradients`."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_mb: float = 8.0, group=None):
        self.group = group
        self.params = [p for p in params if p.requires_grad]
        order = list(reversed(self.params))
        limit = max(1, int(bucket_mb * (1 << 20) / 4))
        self.buckets = [[order[i] for i in idxs] for idxs in plan_buckets([p.numel() for p in order], limit)]
        self.bucket_of = {id(p): bi for bi, ps in enumerate(self.buckets) for p in ps}
        self.ready = [0] * len(self.buckets)
        self.launched = [False] * len(self.buckets)
        self.inflight = []
        self.enabled = False
        self.hooks = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]

    def _active(self) -> bool:
        return self.enabled and dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _hook(self, p):
        if not self._active():
            return
        b = self.bucket_of[id(p)]
        self.ready[b] += 1
        if self.ready[b] == len(self.buckets[b]) and not self.launched[b]:
            self._launch(b)

    def _launch(self, b: int):
        ps = self.buckets[b]
        for p in ps:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        flat = torch.cat([p.grad.reshape(-1) for p in ps])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.inflight.append((work, flat, ps))
        self.launched[b] = True

    def begin(self):
        """Arm the hooks for the backward pass that follows."""
        self.ready = [0] * len(self.buckets)
        self.launched = [False] * len(self.buckets)
        self.inflight = []
        self.enabled = True

    def finish(self, average: bool = True) -> int:
        """After backward(): every rank ends with the (averaged) sum of the gradients.  Returns the elements reduced."""
        total = 0
        if self._active():
            world = dist.get_world_size(self.group)
            for b in range(len(self.buckets)):
                if not self.launched[b]:
                    self._launch(b)
            for work, flat, ps in self.inflight:
                work.wait()
                if average:
                    flat.div_(world)
                off = 0
                for p in ps:
                    n = p.numel()
                    p.grad.copy_(flat[off:off + n].view_as(p.grad))
                    off += n
                total += flat.numel()
        self.inflight = []
        self.enabled = False
        return total

    def close(self):
        for h in self.hooks:
            h.remove()
        self.hooks = []


def shard_weight(n: int, world: int, rank: int) -> float:
    """Weight of this rank's mean loss in the global-batch mean, times `world` (allreduce_gradients divides by it):
    local_items * world / n.  1.0 whenever the batch divides evenly."""
    sl = shard_batch(n, world, rank)
    return (sl.stop - sl.start) * world / float(n)


def shard_batch(n: int, world: int, rank: int) -> slice:
    """Contiguous slice of a global batch of n items for this rank (sizes differ by at most one)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return slice(lo, lo + base + (1 if rank < extra else 0))


# ------------------------------------------------------------------------------------------------------------------
# The adversarial training step of Text2Mel (train/adversarial_wasserstein_gp.py:261-322).
# Generator: spoofsv_b200.models.TTSModel.melSyn in train() mode (highway convs forward / backward in the library's
# kernels).  Discriminator: the reference's melDisc restated in torch ops -- it is 0.3 % of the step's FLOPs and the
# gradient penalty differentiates through its backward pass, so it stays in autograd.
import math

import torch.nn as nn
import torch.nn.functional as F


def _conv1d_mm(x, conv: nn.Conv1d):
    """nn.Conv1d (stride 1, zero padding, dilation) as shifted FP32 matmuls.  cuDNN picks its algorithm (Winograd,
    FFT, ...) by problem size; in the discriminator that made a sharded step differ from the full-batch one by 1e-4
    of the gradient, scaled up by the adversarial weight.  Matmuls are exact FP32 and differentiate twice."""
    w, k = conv.weight, conv.kernel_size[0]
    pad, dil = conv.padding[0], conv.dilation[0]
    if pad:
        x = F.pad(x, (pad, pad))
    T = x.shape[-1] - dil * (k - 1)
    y = None
    for j in range(k):
        term = torch.matmul(w[:, :, j], x[:, :, j * dil: j * dil + T])
        y = term if y is None else y + term
    return y + conv.bias[None, :, None]


class _DiscHighway(nn.Module):
    """models/TTSModel_dropout.py:37-86 (centred taps, dropout 0.05 on the output)."""

    def __init__(self, dimension: int, kernel_size: int, dilation: int):
        super().__init__()
        self.dimension = dimension
        self.conv = nn.Conv1d(dimension, 2 * dimension, kernel_size, padding=dilation * (kernel_size - 1) // 2, dilation=dilation)
        self.ln1 = nn.LayerNorm(dimension)
        self.ln2 = nn.LayerNorm(dimension)
        self.dp = nn.Dropout(p=0.05)

    def forward(self, x):
        h = _conv1d_mm(x, self.conv)
        h1 = self.ln1(h[:, :self.dimension].transpose(1, 2)).transpose(1, 2)
        h2 = self.ln2(h[:, self.dimension:].transpose(1, 2)).transpose(1, 2)
        g = torch.sigmoid(h1)
        return self.dp(g * h2 + (1 - g) * x)


class melDisc(nn.Module):
    """models/discriminator.py:6-42 (and linDisc, :44-80, through the pooling / width arguments): same constructor,
    same state_dict keys, same forward."""

    def __init__(self, freq_bins: int, disc_dim: int, pool1: int = 4, pool2: int = 2, narrow: int = 4):
        super().__init__()
        self.conv1 = nn.Conv1d(freq_bins, disc_dim, 1)
        self.ln1 = nn.LayerNorm(disc_dim)
        self.dp1 = nn.Dropout(p=0.05)
        self.hc = _DiscHighway(disc_dim, 3, 1)
        self.conv2 = nn.Conv1d(disc_dim, 64, 1)
        self.pl1 = nn.AvgPool1d(pool1)
        self.ln2 = nn.LayerNorm(64)
        self.dp2 = nn.Dropout(p=0.05)
        self.conv3 = nn.Conv1d(64, 16, 1)
        self.pl2 = nn.AvgPool1d(pool2)
        self.ln3 = nn.LayerNorm(16)
        self.conv4 = nn.Conv1d(16, narrow, 1)
        self.ln4 = nn.LayerNorm(narrow)
        self.conv5 = nn.Conv1d(narrow, 1, 1)
        self.pl3 = nn.AdaptiveAvgPool1d(1)

    @staticmethod
    def _ln(m, x):
        return m(x.transpose(1, 2)).transpose(1, 2)

    def forward(self, inputs):
        x = self.dp1(self._ln(self.ln1, _conv1d_mm(inputs, self.conv1)))
        x = self.hc(x)
        x = self.dp2(F.leaky_relu(self._ln(self.ln2, self.pl1(_conv1d_mm(x, self.conv2))), 0.05))
        x = self._ln(self.ln3, self.pl2(_conv1d_mm(x, self.conv3)))
        x = self._ln(self.ln4, _conv1d_mm(F.leaky_relu(x, 0.05), self.conv4))
        return self.pl3(_conv1d_mm(F.leaky_relu(x, 0.05), self.conv5))


class linDisc(melDisc):
    """models/discriminator.py:44-80: the SSRN-side discriminator (pooling 8 / 4, 8 channels before the last conv)."""

    def __init__(self, freq_bins: int, disc_dim: int):
        super().__init__(freq_bins, disc_dim, pool1=8, pool2=4, narrow=8)


def guided_attention_mat(max_text_len: int, max_frame_num: int, g: float = 0.2, device=None) -> torch.Tensor:
    """train/adversarial_wasserstein_gp.py guided_attention_mat: W[n, t] = 1 - exp(-(t/T - n/N)^2 / (2 g^2))."""
    n = torch.arange(max_text_len, dtype=torch.float32)[:, None] / max_text_len
    t = torch.arange(max_frame_num, dtype=torch.float32)[None, :] / max_frame_num
    w = 1.0 - torch.exp(-(t - n) ** 2 / (2.0 * g * g))
    return w.to(device) if device is not None else w


import contextlib


@contextlib.contextmanager
def fp32_math():
    """The step is FP32 end to end: torch's convolutions and matmuls would otherwise run the discriminator in TF32
    on this GPU (2e-4 relative noise in every generator gradient through the adversarial term)."""
    conv, mm = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = conv, mm


def generator_step(model, disc, opt_syn, mel_gt, text_id, spk_emb, gaw, cfg, group=None, shard_weight: float = 1.0,
                   reducer: "OverlappedGradReducer | None" = None):
    terms = generator_body(model, disc, opt_syn, mel_gt, text_id, spk_emb, gaw, cfg, group, shard_weight, reducer)
    return generator_terms(terms)


def generator_terms(terms) -> dict:
    """The loss dictionary of a 'G' iteration from the device tensor generator_body returns (one host read)."""
    t_l1, t_bd, t_att, t_disc = (float(v) for v in terms.tolist())
    scale_f = (t_l1 + t_bd + t_att) / abs(t_disc)
    return {"l1": t_l1, "bin_div": t_bd, "att": t_att, "disc": t_disc, "loss": t_l1 + t_bd + t_att + scale_f * t_disc}


def generator_body(model, disc, opt_syn, mel_gt, text_id, spk_emb, gaw, cfg, group=None, shard_weight: float = 1.0,
                   reducer: "OverlappedGradReducer | None" = None):
    """One 'G' iteration (:277-300): teacher-forced forward, L1 + binary divergence + guided attention + adversarial
    term scaled to the size of the other three, backward, gradient allreduce, Adam step.  Returns the four loss terms as a device tensor
    (no host synchronisation anywhere in the body: it can be captured in a CUDA graph, `GraphedIteration`).
    `reducer` (an OverlappedGradReducer over model.parameters()) overlaps the bucket allreduces with the backward pass;
    without it the buckets are reduced after backward.  `shard_weight` = local_items * world / global_items (1.0 for equal shards): the rank's mean losses and their
    gradients are weighted by it, so that the average over ranks is the mean over the GLOBAL batch even when the
    shard sizes differ by one (`shard_weights`)."""
    opt_syn.zero_grad(set_to_none=True)
    spec_inputs = torch.cat((torch.zeros_like(mel_gt[:, :, :1]), mel_gt[:, :, :-1]), dim=-1)
    with fp32_math():
        pred, att = model(spec_inputs, text_id, spk_emb)
        disc_syn = disc(pred)
        loss_l1 = torch.mean(torch.abs(mel_gt - pred))
        loss_bd = torch.mean(-mel_gt * torch.log(pred + 1e-8) - (1 - mel_gt) * torch.log(1 - pred + 1e-8))
        N, T = att.shape[-2], att.shape[-1]
        # the reference pads A with -1 to (MAX_TEXT_LEN, MAX_FRAME_NUM) and masks the padding out again: same value
        loss_att = torch.sum(att * gaw[:N, :T]) / float(att.numel())
        loss_disc = torch.mean(-disc_syn)
        # the adversarial weight is a ratio of GLOBAL-batch loss values (the reference computes it on the gathered
        # DataParallel output): average the four scalars over the ranks first
        terms = torch.stack([loss_l1.detach(), loss_bd.detach(), loss_att.detach(), loss_disc.detach()]) * shard_weight
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(terms, op=dist.ReduceOp.SUM, group=group)
            terms /= dist.get_world_size(group)
        # the weight stays on the device (a detached 0-d tensor): reading the four values back here, as the reference's
        # Python floats do, would stall the host between forward and backward and leave the GPU waiting for the
        # backward kernels to be issued; they are read once, after the optimizer step
        scale = (terms[0] + terms[1] + terms[2]) / terms[3].abs()
        loss = loss_l1 + loss_bd + loss_att + scale * loss_disc
        if reducer is not None:
            reducer.begin()                      # bucket allreduces start from the gradient hooks, under the backward pass
        (loss * shard_weight if shard_weight != 1.0 else loss).backward()
    if reducer is not None:
        reducer.finish()
    else:
        allreduce_gradients(model.parameters(), group=group)
    opt_syn.step()
    return terms


def discriminator_step(model, disc, opt_disc, mel_gt, text_id, spk_emb, cfg, coeff=None, group=None, shard_weight: float = 1.0):
    """One 'D' iteration (:302-322): WGAN-GP.  The generator runs without a graph (its output is detached in the
    reference), the gradient penalty differentiates through the discriminator's backward pass (autograd).
    `coeff` (B,) are the interpolation weights (reference: torch.rand(B), one per utterance)."""
    return wgan_terms(discriminator_body(model, disc, opt_disc, mel_gt, text_id, spk_emb, cfg, coeff, group, shard_weight))


def discriminator_body(model, disc, opt_disc, mel_gt, text_id, spk_emb, cfg, coeff=None, group=None, shard_weight: float = 1.0):
    """discriminator_step without the host read of the loss values: returns the device tensor (gradient penalty,
    discriminator loss)."""
    opt_disc.zero_grad(set_to_none=True)
    spec_inputs = torch.cat((torch.zeros_like(mel_gt[:, :, :1]), mel_gt[:, :, :-1]), dim=-1)
    with torch.no_grad():
        pred, _ = model(spec_inputs, text_id, spk_emb)
    return _wgan_gp_body(disc, opt_disc, mel_gt, pred, cfg, coeff, group, shard_weight)


def wgan_terms(t) -> dict:
    gp, ld = (float(v) for v in t.tolist())
    return {"gp": gp, "wd": -ld, "loss": ld + gp}


def _wgan_gp_step(disc, opt_disc, real, fake, cfg, coeff, group, shard_weight: float = 1.0):
    return wgan_terms(_wgan_gp_body(disc, opt_disc, real, fake, cfg, coeff, group, shard_weight))


def _wgan_gp_body(disc, opt_disc, real, fake, cfg, coeff, group, shard_weight: float = 1.0):
    B = real.shape[0]
    if coeff is None:
        coeff = torch.rand(B, device=real.device)
    c = coeff.to(real.device, real.dtype)[:, None, None]
    mid = (c * real + (1 - c) * fake).requires_grad_(True)
    with fp32_math():
        out_mid = disc(mid)
        grads = torch.autograd.grad(out_mid, mid, torch.ones_like(out_mid), retain_graph=True, create_graph=True)[0]
        loss_gp = torch.mean(cfg["LAMBDA"] * (torch.norm(grads, p=2, dim=(1, 2)) - 1) ** 2)
        (loss_gp * shard_weight if shard_weight != 1.0 else loss_gp).backward()
        loss_d = torch.mean(disc(fake) - disc(real))
        (loss_d * shard_weight if shard_weight != 1.0 else loss_d).backward()
    allreduce_gradients(disc.parameters(), group=group)
    opt_disc.step()
    return torch.stack([loss_gp.detach(), loss_d.detach()])


class GraphedIteration:
    """One training iteration of fixed shapes, captured in a CUDA graph and replayed (single process).

    A generator iteration is ~800 kernel launches and a discriminator iteration several hundred small autograd ops: issued
    one by one the host, not the GPU, sets the pace (15 ms each at the config-5 shape).  `body(*inputs)` must not read
    anything back to the host (generator_body / discriminator_body) and its optimizer must be built with
    `capturable=True`.  The warm-up iterations run on the capture stream (the library's per-stream scratch block and its
    one-time allocations are in place before the capture starts) and ARE training iterations on the first batch.
    `__call__` copies the new batch into the static input tensors, replays, and returns the body's (static) output.
    `parameters`: every parameter whose `.grad` the body writes, including those it does not zero itself (the generator
    iteration also accumulates into the discriminator's gradients).  Their `.grad` is dropped before the capture so that
    the graph allocates them in its own pool: a gradient tensor left over from an eager iteration would be captured by
    address and may be returned to the driver later (the next capture empties the allocator's cache)."""

    def __init__(self, body, inputs, warmup: int = 3, parameters=None):
        self.inputs = [t.clone() for t in inputs]
        self.stream = torch.cuda.Stream()
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            for _ in range(warmup):
                body(*self.inputs)
        torch.cuda.current_stream().wait_stream(self.stream)
        torch.cuda.synchronize()
        for p in (parameters or ()):
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            self.out = body(*self.inputs)

    def __call__(self, *inputs):
        for dst, src in zip(self.inputs, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src)
        self.graph.replay()
        return self.out


def ssrn_generator_step(model, disc, opt_syn, mel_gt, lin_gt, cfg=None, group=None, shard_weight: float = 1.0):
    """One 'G' iteration of `train_ssrn` (train/adversarial_wasserstein_gp.py:324-338): SSRN forward, L1 + binary
    divergence + the adversarial term scaled to their size, backward, gradient allreduce, optimizer step."""
    opt_syn.zero_grad(set_to_none=True)
    with fp32_math():
        pred = model(mel_gt)
        disc_syn = disc(pred)
        loss_l1 = torch.mean(torch.abs(lin_gt - pred))
        loss_bd = torch.mean(-lin_gt * torch.log(pred + 1e-8) - (1 - lin_gt) * torch.log(1 - pred + 1e-8))
        loss_disc = torch.mean(-disc_syn)
        terms = torch.stack([loss_l1.detach(), loss_bd.detach(), loss_disc.detach()]) * shard_weight
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(terms, op=dist.ReduceOp.SUM, group=group)
            terms /= dist.get_world_size(group)
        scale = (terms[0] + terms[1]) / terms[2].abs()           # on the device, see generator_step
        ((loss_l1 + loss_bd + scale * loss_disc) * shard_weight).backward()
    allreduce_gradients(model.parameters(), group=group)
    opt_syn.step()
    t_l1, t_bd, t_disc = (float(v) for v in terms.tolist())
    scale_f = (t_l1 + t_bd) / abs(t_disc)
    return {"l1": t_l1, "bin_div": t_bd, "disc": t_disc, "loss": t_l1 + t_bd + scale_f * t_disc}


def ssrn_discriminator_step(model, disc, opt_disc, mel_gt, lin_gt, cfg, coeff=None, group=None, shard_weight: float = 1.0):
    """One 'D' iteration of `train_ssrn` (:340-360): WGAN-GP on linear spectrograms."""
    opt_disc.zero_grad(set_to_none=True)
    with torch.no_grad():
        pred = model(mel_gt)
    return _wgan_gp_step(disc, opt_disc, lin_gt, pred, cfg, coeff, group, shard_weight)


def ordinary_step(model, opt, batch, gaw, text2mel: bool, group=None, shard_weight: float = 1.0):
    """One iteration of the non-adversarial trainer (train/ordinary.py:219-256): L1 + binary divergence
    (+ guided attention for Text2Mel), backward, gradient allreduce, optimizer step."""
    opt.zero_grad(set_to_none=True)
    with fp32_math():
        if text2mel:
            mel_gt, text_id, spk_emb = batch
            spec_inputs = torch.cat((torch.zeros_like(mel_gt[:, :, :1]), mel_gt[:, :, :-1]), dim=-1)
            pred, att = model(spec_inputs, text_id, spk_emb)
            target = mel_gt
        else:
            mel_gt, target = batch
            pred, att = model(mel_gt), None
        loss_l1 = torch.mean(torch.abs(target - pred))
        loss_bd = torch.mean(-target * torch.log(pred + 1e-8) - (1 - target) * torch.log(1 - pred + 1e-8))
        loss = loss_l1 + loss_bd
        out = {"l1": loss_l1.item(), "bin_div": loss_bd.item()}
        if att is not None:
            loss_att = torch.sum(att * gaw[:att.shape[-2], :att.shape[-1]]) / float(att.numel())
            loss = loss + loss_att
            out["att"] = loss_att.item()
        (loss * shard_weight if shard_weight != 1.0 else loss).backward()
    allreduce_gradients(model.parameters(), group=group)
    opt.step()
    out["loss"] = loss.item()
    return out
