"""CLI-compatible entry point for the reference's main.py.

    python -m spoofsv_b200.main {train_text2mel,train_ssrn,synthesize} -C config.json -T <time tag>
                                [-P conditional] [-R checkpoint] [--adversarial] [--save_spectrogram]
    python -m torch.distributed.run --nproc-per-node N -m spoofsv_b200.main train_text2mel --adversarial ...

Same positional step, same flags and the same config.json keys as main.py:10-16, the same output locations
(`SRC_ROOT_DIR/samples/<T>/S<k>_B<i>.wav`, `SRC_ROOT_DIR/checkpoints/<pattern>/adversarial/<T>/...tar.pth`) and the
same checkpoint dictionary keys (train/adversarial_wasserstein_gp.py:401-434), so a run can be resumed by either
implementation.  What is different: synthesis uses the incremental decode kernel instead of re-encoding the mel
prefix every frame; the waveform stage is batched on the GPU; training is one process per GPU with a gradient
allreduce (torchrun) instead of nn.DataParallel; attention plots (matplotlib) are not produced.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from typing import Optional

import numpy as np
import torch


def build_parser() -> argparse.ArgumentParser:
    ps = argparse.ArgumentParser(description="Adversarial Conditional Text-to-speech")
    ps.add_argument("step", choices=["train_text2mel", "train_ssrn", "synthesize"], metavar="s")
    ps.add_argument("-P", "--pattern", choices=["universal", "conditional", "ubm-finetune"], default="conditional", metavar="m")
    ps.add_argument("-R", "--resume", type=str, default=None, metavar="checkpoint")
    ps.add_argument("-C", "--configuration", type=str, default=None)
    ps.add_argument("--adversarial", action="store_true")
    ps.add_argument("--save_spectrogram", action="store_true")
    ps.add_argument("-T", "--current_time", type=str, required=True, metavar="T")
    # extensions
    ps.add_argument("--max_iterations", type=int, default=None, help="stop after this many global iterations")
    ps.add_argument("--random_init", type=int, default=None, metavar="SEED",
                    help="synthesize: random-init weights instead of the INFERENCE_* checkpoints")
    ps.add_argument("--gl_iters", type=int, default=64)
    return ps


def _models(cfg, pattern):
    from .models.TTSModel import SSRN, melSyn
    if pattern != "conditional":
        raise NotImplementedError("spoofsv_b200 implements the speaker-conditioned pattern ('conditional')")
    m1 = melSyn(vocab_len=len(cfg["VOCABULARY"]) - 1, condition=True, spkemb_dim=cfg["SPK_EMB_DIM"],
                textemb_dim=cfg["TEXT_EMB_DIM"], freq_bins=cfg["COARSE_MELSPEC"]["FREQ_BINS"], hidden_dim=cfg["HIDDEN_DIM"])
    m2 = SSRN(freq_bins=cfg["COARSE_MELSPEC"]["FREQ_BINS"], output_bins=1 + cfg["STFT"]["FFT_LENGTH"] // 2,
              ssrn_dim=cfg["SSRN_DIM"])
    return m1, m2


def _spectral_losses(pred, gt):
    l1 = torch.mean(torch.abs(gt - pred))
    bd = torch.mean(-gt * torch.log(pred + 1e-8) - (1 - gt) * torch.log(1 - pred + 1e-8))
    return l1, bd


def synthesize(pattern: str, cfg: dict, spec_dir: Optional[str], current_time: str, random_init: Optional[int] = None,
               gl_iters: int = 64) -> dict:
    """synthesize.py:41-147: every batch of 8 items of the 'synthesize' list is decoded for as many frames as its
    longest ground-truth mel, scored against the ground truth, run through SSRN and the waveform stage."""
    from torch.utils.data import DataLoader
    from . import vocoder
    from .data import collate_pad_4, dataset
    from .train import guided_attention_mat
    sample_dir = cfg["SRC_ROOT_DIR"] + "samples/" + current_time + "/"
    os.makedirs(sample_dir, exist_ok=True)
    m1, m2 = _models(cfg, pattern)
    if random_init is not None:
        torch.manual_seed(random_init)
        m1, m2 = _models(cfg, pattern)
    else:
        m1.load_state_dict(torch.load(cfg["INFERENCE_TEXT2MEL_MODEL"], map_location="cpu")["model_state_dict"])
        m2.load_state_dict(torch.load(cfg["INFERENCE_SSRN_MODEL"], map_location="cpu")["model_state_dict"])
    m1, m2 = m1.cuda().eval(), m2.cuda().eval()
    loader = DataLoader(dataset(cfg=cfg, mode="synthesize", pattern=pattern, step="synthesize", spec_dir=spec_dir),
                        batch_size=8, shuffle=False, num_workers=0, collate_fn=collate_pad_4)
    gaw = guided_attention_mat(cfg["MAX_TEXT_LEN"], cfg["MAX_FRAME_NUM"], device="cuda")
    tot_t2m = tot_ssrn = 0.0
    n_batches = n_wavs = 0
    with torch.no_grad():
        for i, sp in enumerate(loader):
            mel_gt, text_id, spk_emb, lin_gt = (sp[k].cuda() for k in ("data_0", "data_1", "data_2", "data_3"))
            B, _, T = mel_gt.shape
            Y, A, _, _, _ = m1.synthesize(text_id, spk_emb, T)                      # the AR loop of :103-109
            l1, bd = _spectral_losses(Y, mel_gt)
            att = torch.sum(A * gaw[:A.shape[-2], :A.shape[-1]]) / float(A.numel())
            print("syn set text2mel loss: {} {} {} {}".format(l1.item(), bd.item(), att.item(), (l1 + bd + att).item()))
            tot_t2m += (l1 + bd + att).item()
            lin = m2(Y)
            l1s, bds = _spectral_losses(lin, lin_gt[:, :, :lin.shape[-1]] if lin_gt.shape[-1] >= lin.shape[-1] else
                                        torch.nn.functional.pad(lin_gt, (0, lin.shape[-1] - lin_gt.shape[-1])))
            print("syn set ssrn loss: {} {} {}".format(l1s.item(), bds.item(), (l1s + bds).item()))
            tot_ssrn += (l1s + bds).item()
            # synthesize.py:134-147 keeps the whole signal (no trim, no 9 s cap): Griffin-Lim, de-emphasis, peak 0.75;
            # LOG_FEATURE (:134-136): dB mapping instead of the max-normalisation, signal written as is
            x = lin.to(torch.float32)
            log_feature = bool(cfg.get("LOG_FEATURE", False))
            power = cfg["NORM_POWER"]["RECONSTRUCTION"] / cfg["NORM_POWER"]["ANALYSIS"]
            if log_feature:
                spec = torch.pow(10.0, 0.05 * (x * cfg["MAX_DB"] - cfg["MAX_DB"] + cfg["REF_DB"])).pow(power)
            else:
                spec = (x / x.amax(dim=(1, 2), keepdim=True)).pow(power)
            sig = vocoder.deemphasis(vocoder.griffin_lim(spec, gl_iters, cfg["STFT"]["HOP_LENGTH"], cfg["STFT"]["FFT_LENGTH"]),
                                     cfg["PREEMPH"])
            sig = (sig if log_feature else sig / sig.amax(dim=1, keepdim=True) * 0.75).cpu().numpy()
            for k in range(B):
                vocoder.write_wav(sample_dir + "S{}_B{}.wav".format(k + 1, i + 1), sig[k], cfg["SAMPLING_RATE"])
                n_wavs += 1
            n_batches += 1
    out = {"batches": n_batches, "wavs": n_wavs, "text2mel_loss": tot_t2m / max(n_batches, 1),
           "ssrn_loss": tot_ssrn / max(n_batches, 1), "sample_dir": sample_dir}
    print(json.dumps(out), flush=True)
    return out


def _init_weights(layer):
    """train/adversarial_wasserstein_gp.py:15-18."""
    if hasattr(layer, "weight") and layer.weight is not None and len(layer.weight.shape) > 1:
        torch.nn.init.kaiming_normal_(layer.weight, nonlinearity="relu")


def _validate(loader, gaw, cfg, model, train_step):
    """:65-107 (validation loss over the loader, free-running decode for Text2Mel)."""
    tot, n = 0.0, 0
    with torch.no_grad():
        for sp in loader:
            mel_gt = sp["data_0"].cuda()
            if train_step == "train_text2mel":
                Y, A, _, _, _ = model.synthesize(sp["data_1"].cuda(), sp["data_2"].cuda(), mel_gt.shape[-1])
                l1, bd = _spectral_losses(Y, mel_gt)
                loss = l1 + bd + torch.sum(A * gaw[:A.shape[-2], :A.shape[-1]]) / float(A.numel())
            else:
                l1, bd = _spectral_losses(model(mel_gt), sp["data_1"].cuda())
                loss = l1 + bd
            tot += loss.item()
            n += 1
    return tot / max(n, 1)


def adversarial_train(train_step: str, train_pattern: str, cfg: dict, spec_dir: Optional[str], resume_checkpoints: Optional[str],
                      current_time: str, max_iterations: Optional[int] = None) -> dict:
    """train/adversarial_wasserstein_gp.py:142-447: RATIO discriminator iterations per generator iteration, Adam on
    both, validation and checkpoints every VAL_EVERY_ITER iterations.  Under torchrun every rank takes its slice of
    each batch and the gradients are all-reduced."""
    import torch.distributed as dist
    from torch.utils.data import DataLoader
    from . import train as TR
    from .data import collate_pad_2, collate_pad_3, dataset
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl")
    save_dir = cfg["SRC_ROOT_DIR"] + "checkpoints/" + train_pattern + "/adversarial/" + current_time
    if rank == 0:
        os.makedirs(save_dir, exist_ok=True)
    text2mel = train_step == "train_text2mel"
    m1, m2 = _models(cfg, train_pattern)
    model = m1 if text2mel else m2
    disc = (TR.melDisc(cfg["COARSE_MELSPEC"]["FREQ_BINS"], cfg["DISC_DIM"]) if text2mel
            else TR.linDisc(1 + cfg["STFT"]["FFT_LENGTH"] // 2, cfg["DISC_DIM"]))
    logs = {"wd_log": [], "loss_train_log_syn": [], "loss_train_log_syn_onlyfromD": [], "loss_train_log_disc": [], "loss_val_log": []}
    epoch = iteration = 0
    adam = cfg["ADAM"]
    mk_opt = lambda m: torch.optim.Adam(m.parameters(), adam["ALPHA"], (adam["BETA_1"], adam["BETA_2"]), adam["EPSILON"])
    if resume_checkpoints is None:
        torch.manual_seed(0)                                # same initial weights on every rank
        model.apply(_init_weights)
        disc.apply(_init_weights)
        model, disc = model.cuda(), disc.cuda()
        opt_syn, opt_disc = mk_opt(model), mk_opt(disc)
    else:
        ck = torch.load(resume_checkpoints, map_location="cpu")
        epoch, iteration = ck["epoch"], ck["iteration"]
        model.load_state_dict(ck["model_state_dict"])
        disc.load_state_dict(ck["disc_state_dict"])
        model, disc = model.cuda(), disc.cuda()
        opt_syn, opt_disc = mk_opt(model), mk_opt(disc)
        opt_syn.load_state_dict(ck["opt_state_dict_syn"])
        opt_disc.load_state_dict(ck["opt_state_dict_disc"])
        for k in logs:
            logs[k] = ck.get(k, [])
    model.train()
    disc.train()
    collate = collate_pad_3 if text2mel else collate_pad_2
    mk_loader = lambda mode, bs, shuffle: DataLoader(
        dataset(cfg=cfg, mode=mode, pattern=train_pattern, step=train_step, spec_dir=spec_dir),
        batch_size=bs, shuffle=shuffle, num_workers=0, collate_fn=collate, generator=torch.Generator().manual_seed(1234 + epoch))
    train_loader, val_loader = mk_loader("train", cfg["BATCH_SIZE"], True), mk_loader("validate", 8, False)
    gaw = TR.guided_attention_mat(cfg["MAX_TEXT_LEN"], cfg["MAX_FRAME_NUM"], device="cuda")
    # generator iterations of Text2Mel: bucket allreduces launched from gradient hooks, overlapped with backward
    reducer = TR.OverlappedGradReducer(model.parameters()) if (world > 1 and text2mel) else None
    done = False
    t_start = time.perf_counter()
    while epoch < cfg["MAX_EPOCHS"] and not done:
        for sp in train_loader:                              # every rank draws the same batch and keeps its slice
            n_items = sp["data_0"].shape[0]
            if n_items < world:
                continue                                     # decided from the GLOBAL batch: every rank skips together, so no
                                                             # rank misses a collective or falls out of the G / D schedule
            sl = TR.shard_batch(n_items, world, rank)
            sw = TR.shard_weight(n_items, world, rank)       # uneven shards: weight by local / global items
            target = "D" if iteration % (cfg["RATIO"] + 1) else "G"
            a = sp["data_0"][sl].cuda()
            if text2mel:
                ids, spk = sp["data_1"][sl].cuda(), sp["data_2"][sl].cuda()
                if target == "G":
                    t = TR.generator_step(model, disc, opt_syn, a, ids, spk, gaw, cfg, shard_weight=sw, reducer=reducer)
                else:
                    t = TR.discriminator_step(model, disc, opt_disc, a, ids, spk, cfg, shard_weight=sw)
            else:
                lin_gt = sp["data_1"][sl].cuda()
                if target == "G":
                    t = TR.ssrn_generator_step(model, disc, opt_syn, a, lin_gt, cfg, shard_weight=sw)
                else:
                    t = TR.ssrn_discriminator_step(model, disc, opt_disc, a, lin_gt, cfg, shard_weight=sw)
            if target == "G":
                logs["loss_train_log_syn"].append(t["loss"])
                logs["loss_train_log_syn_onlyfromD"].append(t["disc"])
            else:
                logs["loss_train_log_disc"].append(t["loss"])
                logs["wd_log"].append(t["wd"])
            if rank == 0:
                print("Global iteration {} training {}: {}".format(iteration + 1, target, json.dumps(t)), flush=True)
            if iteration % cfg["VAL_EVERY_ITER"] == 0 and iteration > 0:
                model.eval()
                loss_val = _validate(val_loader, gaw, cfg, model, train_step)
                model.train()
                logs["loss_val_log"].append(loss_val)
                if rank == 0:
                    ck = {"epoch": epoch + 1, "iteration": iteration + 1, "model_state_dict": model.state_dict(),
                          "disc_state_dict": disc.state_dict(), "opt_state_dict_syn": opt_syn.state_dict(),
                          "opt_state_dict_disc": opt_disc.state_dict(), **logs}
                    if logs["loss_val_log"].index(min(logs["loss_val_log"])) == len(logs["loss_val_log"]) - 1:
                        torch.save(ck, save_dir + "/{}_best_model.tar.pth".format(train_step[6:]))
                    torch.save(ck, save_dir + "/{}_iteration_{}.tar.pth".format(train_step[6:], iteration + 1))
                    print("validation loss {} at iteration {}; checkpoint saved in {}".format(loss_val, iteration + 1, save_dir), flush=True)
            iteration += 1
            if max_iterations is not None and iteration >= max_iterations:
                done = True
                break
        epoch += 1
    out = {"iterations": iteration, "epochs": epoch, "seconds": time.perf_counter() - t_start, "save_dir": save_dir,
           "last_G": logs["loss_train_log_syn"][-1] if logs["loss_train_log_syn"] else None,
           "last_D": logs["loss_train_log_disc"][-1] if logs["loss_train_log_disc"] else None}
    if rank == 0:
        print(json.dumps(out), flush=True)
    return out


def ordinary_train(train_step: str, train_pattern: str, cfg: dict, spec_dir: Optional[str], resume_checkpoints: Optional[str],
                   current_time: str, max_iterations: Optional[int] = None) -> dict:
    """train/ordinary.py:130-292: the trainer without a discriminator (checkpoints under .../not_adversarial/<T>/ with
    the keys epoch, iteration, model_state_dict, optimizer_state_dict, loss_val_log)."""
    import torch.distributed as dist
    from torch.utils.data import DataLoader
    from . import train as TR
    from .data import collate_pad_2, collate_pad_3, dataset
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl")
    save_dir = cfg["SRC_ROOT_DIR"] + "checkpoints/" + train_pattern + "/not_adversarial/" + current_time
    if rank == 0:
        os.makedirs(save_dir, exist_ok=True)
    text2mel = train_step == "train_text2mel"
    m1, m2 = _models(cfg, train_pattern)
    model = m1 if text2mel else m2
    adam = cfg["ADAM"]
    mk_opt = lambda m: torch.optim.Adam(m.parameters(), adam["ALPHA"], (adam["BETA_1"], adam["BETA_2"]), adam["EPSILON"])
    epoch = iteration = 0
    loss_val_log = []
    if resume_checkpoints is None:
        torch.manual_seed(0)
        model.apply(_init_weights)
        model = model.cuda()
        opt = mk_opt(model)
    else:
        ck = torch.load(resume_checkpoints, map_location="cpu")
        epoch, iteration, loss_val_log = ck["epoch"], ck["iteration"], ck.get("loss_val_log", [])
        model.load_state_dict(ck["model_state_dict"])
        model = model.cuda()
        opt = mk_opt(model)
        opt.load_state_dict(ck["optimizer_state_dict"])
    model.train()
    collate = collate_pad_3 if text2mel else collate_pad_2
    mk_loader = lambda mode, bs, shuffle: DataLoader(
        dataset(cfg=cfg, mode=mode, pattern=train_pattern, step=train_step, spec_dir=spec_dir),
        batch_size=bs, shuffle=shuffle, num_workers=0, collate_fn=collate, generator=torch.Generator().manual_seed(1234 + epoch))
    train_loader, val_loader = mk_loader("train", cfg["BATCH_SIZE"], True), mk_loader("validate", 8, False)
    gaw = TR.guided_attention_mat(cfg["MAX_TEXT_LEN"], cfg["MAX_FRAME_NUM"], device="cuda")
    done, last = False, None
    while epoch < cfg["MAX_EPOCHS"] and not done:
        for sp in train_loader:
            n_items = sp["data_0"].shape[0]
            if n_items < world:
                continue                                     # every rank skips together (see adversarial_train)
            sl = TR.shard_batch(n_items, world, rank)
            keys = ("data_0", "data_1", "data_2") if text2mel else ("data_0", "data_1")
            last = TR.ordinary_step(model, opt, tuple(sp[k][sl].cuda() for k in keys), gaw, text2mel,
                                    shard_weight=TR.shard_weight(n_items, world, rank))
            if rank == 0:
                print("global iteration {}: {}".format(iteration + 1, json.dumps(last)), flush=True)
            if iteration % cfg["VAL_EVERY_ITER"] == 0 and iteration > 0:
                model.eval()
                loss_val_log.append(_validate(val_loader, gaw, cfg, model, train_step))
                model.train()
                if rank == 0:
                    ck = {"epoch": epoch + 1, "iteration": iteration + 1, "model_state_dict": model.state_dict(),
                          "optimizer_state_dict": opt.state_dict(), "loss_val_log": loss_val_log}
                    if loss_val_log.index(min(loss_val_log)) == len(loss_val_log) - 1:
                        torch.save(ck, save_dir + "/{}_best_model.tar.pth".format(train_step[6:]))
                    torch.save(ck, save_dir + "/{}_iteration_{}.tar.pth".format(train_step[6:], iteration + 1))
            iteration += 1
            if max_iterations is not None and iteration >= max_iterations:
                done = True
                break
        epoch += 1
    out = {"iterations": iteration, "epochs": epoch, "save_dir": save_dir, "last": last}
    if rank == 0:
        print(json.dumps(out), flush=True)
    return out


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    if args.configuration is None:
        print("main: -C/--configuration is required", file=sys.stderr)
        return 2
    with open(args.configuration, "r") as f:
        cfg = json.load(f)
    if not torch.cuda.is_available():
        raise RuntimeError("spoofsv_b200.main: no CUDA device (there is no CPU path)")
    spec_dir = None
    if args.save_spectrogram:
        spec_dir = cfg["SRC_ROOT_DIR"] + "spec/"
        os.makedirs(spec_dir, exist_ok=True)
    if args.step in ("train_text2mel", "train_ssrn"):
        trainer = adversarial_train if args.adversarial else ordinary_train
        trainer(args.step, args.pattern, cfg, spec_dir, args.resume, args.current_time, args.max_iterations)
    else:
        synthesize(args.pattern, cfg, spec_dir, args.current_time, args.random_init, args.gl_iters)
    return 0


if __name__ == "__main__":
    sys.exit(main())
