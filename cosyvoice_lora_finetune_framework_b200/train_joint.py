"""Flow / joint LoRA training entry point (CLI mirror of the reference's train_joint.py: `--mode
--resume --epochs --batch-size --lr`, same checkpoint and merged-weight file names) with its own
loop instead of PyTorch-Lightning, plus `--devices N` data parallelism (one process per GPU under
torchrun; one NCCL allreduce of the flat LoRA-gradient bucket per optimiser step).

  python -m cosyvoice_lora_finetune_framework_b200.train_joint --mode flow_only [--synthetic 64]
  torchrun --nproc-per-node 8 -m cosyvoice_lora_finetune_framework_b200.train_joint --mode flow_only

Semantics kept from the reference trainer (train_joint.py:105-226, 312-384): AdamW(lr, wd 0.01,
betas 0.9/0.999), linear warm-up -> cosine decay per optimiser step, gradient clip 1.0, gradient
accumulation, `joint_{mode}_last.ckpt` with `state_dict` keys prefixed `model.flow.` / `model.llm.`,
stop when flow_loss <= 0.3 (LossThresholdCallback), `flow_merged_{mode}.pt` written at the end.
Batches follow the reference's collate schema (dataset.py:549-596): speech_token[_len],
speech_feat[_len], embedding (+ optional text_token[_len], cross_sample_mel[_len]).
"""
import argparse
import os
import time

import torch
import torch.distributed as dist

from .config import DATA_DIR, JOINT_TRAINING_CONFIG, OUTPUT_DIR, PRETRAINED_MODEL_DIR, TRAIN_CONFIG
from .llm_flow_model import build_joint_model, get_joint_merged_state_dict
from .trainer import FlowLoRATrainer


def synthetic_batches(n_batches, batch_size, max_feat_len=250, seed=0):
    """Batches in the reference's collate schema with random content (for smoke runs / benchmarks)."""
    g = torch.Generator().manual_seed(seed)
    for _ in range(n_batches):
        feat_len = torch.randint(max_feat_len // 2, max_feat_len + 1, (batch_size,), generator=g)
        tok_len = (feat_len.float() * 256 * 50 / 22050).long().clamp(min=1)
        T, N = int(feat_len.max()), int(tok_len.max())
        feat = torch.full((batch_size, T, 80), -11.5)
        tok = torch.zeros(batch_size, N, dtype=torch.long)
        for i in range(batch_size):
            feat[i, : feat_len[i]] = torch.randn(int(feat_len[i]), 80, generator=g) * 2 - 6
            tok[i, : tok_len[i]] = torch.randint(0, 4096, (int(tok_len[i]),), generator=g)
        yield {'speech_token': tok, 'speech_token_len': tok_len, 'speech_feat': feat, 'speech_feat_len': feat_len,
               'embedding': torch.randn(batch_size, 192, generator=g)}


def _data_iter(args, batch_size, rank, world):
    if args.synthetic:
        return list(synthetic_batches(args.synthetic, batch_size, JOINT_TRAINING_CONFIG.get('max_feat_len', 250),
                                      seed=rank))
    try:   # the reference's parquet dataset, when the caller has it on the path (out of scope here)
        from dataset import FlowFinetuneDataset, collate_fn
    except ImportError as e:
        raise SystemExit("no dataset module on the path (%s); pass --synthetic N for a smoke run" % e)
    from torch.utils.data import DataLoader
    from torch.utils.data.distributed import DistributedSampler
    ds = FlowFinetuneDataset(data_dir=DATA_DIR)
    sampler = DistributedSampler(ds, world, rank, shuffle=True, drop_last=True) if world > 1 else None
    return DataLoader(ds, batch_size=batch_size, shuffle=sampler is None, sampler=sampler, num_workers=0,
                      collate_fn=collate_fn, pin_memory=True, drop_last=True)


def save_checkpoint(model, trainer, path, epoch, loss):
    """Lightning-layout checkpoint (`state_dict` keys prefixed `model.`) plus the optimiser state needed for an exact
    resume: Adam moments of both flat buckets, the step counter (= schedule position) and the loss scale."""
    sd = {'model.' + k: v.detach().cpu() for k, v in model.state_dict().items()}
    opt = trainer.state_dict()
    torch.save({'state_dict': sd, 'epoch': epoch, 'train_loss': loss, 'global_step': opt['step'], 'optimizer': opt}, path)


def main(argv=None):
    ap = argparse.ArgumentParser(description='LLM + Flow joint LoRA training (B200-native flow path)')
    ap.add_argument('--mode', type=str, default='flow_only', choices=['joint', 'llm_only', 'flow_only'])
    ap.add_argument('--resume', type=str, default=None)
    ap.add_argument('--epochs', type=int, default=None)
    ap.add_argument('--batch-size', type=int, default=None)
    ap.add_argument('--lr', type=float, default=None)
    ap.add_argument('--synthetic', type=int, default=0, help='train on N synthetic batches per epoch')
    ap.add_argument('--dtype', default='fp16', choices=['fp16', 'bf16'])
    ap.add_argument('--output-dir', default=OUTPUT_DIR)
    args = ap.parse_args(argv)
    if args.mode != 'flow_only':
        raise SystemExit("this build accelerates the flow path; --mode %s needs the upstream LLM (SURVEY 8f-4)" % args.mode)

    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    cfg = JOINT_TRAINING_CONFIG
    epochs = args.epochs or cfg.get('max_epochs', 50)
    batch_size = args.batch_size or cfg.get('batch_size', 1)
    lr = args.lr or cfg.get('learning_rate', 5e-5)
    accumulate = cfg.get('accumulate_grad_batches', 8)

    # flow LoRA exactly as configured by the reference (config.py:207-216: r, alpha, lora_dropout, target list): the
    # estimator's attn1 q/k/v run on the CUDA path (dropout included), the Conformer encoder's linear_q/k/v, w_1, w_2
    # are host-side modules trained through dL/dmu of the estimator backward.
    flow_lora = dict(cfg.get('flow_lora', {}))
    model = build_joint_model(PRETRAINED_MODEL_DIR, str(device), 'flow_only', None, flow_lora)
    model.flow.decoder.estimator.cvflow_dtype = torch.float16 if args.dtype == 'fp16' else torch.bfloat16
    model.flow.encoder_autocast = model.flow.decoder.estimator.cvflow_dtype      # '16-mixed' (config.py:76) for the host-side encoder
    upstream = [p for n, p in model.flow.named_parameters() if p.requires_grad and not n.startswith('decoder.estimator.')]
    resume, start_epoch = None, 0
    if args.resume:      # Lightning's ckpt_path semantics: parameters, optimiser moments, schedule position, epoch
        resume = torch.load(args.resume, map_location='cpu', weights_only=False)
        state = resume.get('state_dict', {})
        model.load_state_dict({k[len('model.'):]: v for k, v in state.items() if k.startswith('model.')}, strict=False)
        start_epoch = int(resume.get('epoch', -1)) + 1
    data = _data_iter(args, batch_size, rank, world)
    steps_per_epoch = max(1, len(data) // accumulate)
    trainer = FlowLoRATrainer(model.flow.decoder, lr=lr, weight_decay=TRAIN_CONFIG.get('weight_decay', 0.01),
                              max_grad_norm=TRAIN_CONFIG.get('gradient_clip_val', 1.0),
                              warmup_steps=TRAIN_CONFIG.get('warmup_steps', 50), total_steps=epochs * steps_per_epoch,
                              min_lr=TRAIN_CONFIG.get('min_learning_rate', 1e-6), accumulate=accumulate,
                              extra_params=upstream)
    if resume is not None and isinstance(resume.get('optimizer'), dict) and 'm' in resume['optimizer']:
        trainer.load_state_dict(resume['optimizer'])
    model.train()
    os.makedirs(args.output_dir, exist_ok=True)
    for epoch in range(start_epoch, epochs):
        t0, losses = time.time(), []
        for batch in data:
            out = model(batch, device)
            (out['loss'] / accumulate).backward()
            trainer.micro += 1
            if trainer.micro >= accumulate:
                trainer.optimizer_step()
            losses.append(out['flow_loss'].detach())
        if trainer.micro > 0:      # Lightning steps the optimiser on the last batch of an epoch: no carry-over
            trainer.optimizer_step()
        skipped = trainer.poll_overflow()      # fp16: steps skipped on a non-finite gradient norm -> loss scale halved
        mean_t = torch.stack(losses).mean()
        if world > 1:             # every rank must take the same early-stop decision (else the others hang in all_reduce)
            dist.all_reduce(mean_t, op=dist.ReduceOp.AVG)
        mean = float(mean_t)
        if rank == 0:
            print(f"epoch {epoch}: flow_loss {mean:.4f} lr {trainer.current_lr():.2e} ({time.time() - t0:.1f}s)"
                  + (f" [{skipped} step(s) skipped on overflow, loss scale now {trainer.ne.loss_scale:g}]" if skipped else ""))
            save_checkpoint(model, trainer, os.path.join(args.output_dir, f'joint_{args.mode}_last.ckpt'), epoch, mean)
        if mean <= 0.3:   # LossThresholdCallback(flow <= 0.3), reference train_joint.py:336-340
            break
    if rank == 0:
        merged = get_joint_merged_state_dict(model)
        if 'flow' in merged:
            torch.save(merged['flow'], os.path.join(args.output_dir, f'flow_merged_{args.mode}.pt'))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
