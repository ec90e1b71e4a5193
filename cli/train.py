"""cli/train.py -- drop-in for the reference entry point (cli/train.py:16-89) for the unit-LM recipe, driving the
sm_100a train step instead of HF Trainer.

    torchrun --nproc-per-node 8 cli/train.py data.train_path=<tokens.jsonl> data.val_path=<tokens.jsonl> \
        model=slam model.tlm_type=b200 training_args.output_dir=<dir> [+training_args.max_steps=N]

Data contract of `init_dataset` (slamkit/data/hf_dataset.py:91-118): tokenise `audio_repr`, chunk to `context_len`,
right-pad with 0, labels = input_ids with pad -> -100.  Schedule / clip / AdamW as config/training_args/default.yaml.
`model.config_args.twist_init=false` (or an unreachable base model) starts from seeded random weights."""
import glob
import json
import logging
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from slamkit_b200.config import load_config, require, to_container  # noqa: E402
from slamkit_b200.tokeniser import B200UnitTokeniser  # noqa: E402

logger = logging.getLogger(__name__)


def parse_run_time(rt) -> int:
    """RunTimeStopperCallback (slamkit/trainer/callbacks.py:15-27): "D-HH:MM:SS" or seconds."""
    if isinstance(rt, int):
        return rt
    days = 0
    if "-" in rt:
        d, rt = rt.split("-")
        days = int(d)
    h, m, s = rt.split(":")
    return days * 86400 + int(h) * 3600 + int(m) * 60 + int(s)


def load_chunks(pattern: str, tok: B200UnitTokeniser, context_len: int, min_len=None, max_len=None):
    out = []
    for path in sorted(glob.glob(pattern)):
        for line in open(path):
            ids = tok.prepare_sample(json.loads(line))["input_ids"]
            if max_len and len(ids) > max_len:
                continue
            for i in range(0, len(ids), context_len):          # chunk_texts: keep the remainder, no extra specials
                ch = ids[i:i + context_len]
                if min_len and len(ch) < min_len:
                    continue
                out.append(ch)
    return out


def collate(chunks, pad_id: int = 0):
    """DataCollatorForLanguageModeling(mlm=False): right-pad, labels = ids with pad -> -100."""
    n = max(len(c) for c in chunks)
    ids = torch.full((len(chunks), n), pad_id, dtype=torch.int64)
    for i, c in enumerate(chunks):
        ids[i, :len(c)] = torch.tensor(c)
    labels = ids.clone()
    labels[ids == pad_id] = -100
    return {"input_ids": ids, "labels": labels}


def main(argv=None):
    cfg = load_config("train", argv if argv is not None else sys.argv[1:])
    require(cfg, "data.train_path", "data.val_path")
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    ta = cfg.training_args
    if cfg.data.get("packing", False):
        raise ValueError("Packing is only supported with flash_attention_2 model")      # cli/train.py:43-45
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    p = cfg.tokeniser.params
    tok = B200UnitTokeniser(None, dedup=p.dedup, bos_eos_token_id=p.get("bos_eos_token_id", 1), pad_token_id=p.pad_token_id,
                            num_units=p.get("num_units") or cfg.tokeniser.feature_extractor.num_units, load_fe=False)
    if cfg.model.config_args.vocab_size == -1:
        cfg.model.config_args.vocab_size = len(tok)
    ctx = cfg.model.context_len
    train = load_chunks(cfg.data.train_path, tok, ctx, cfg.data.get("chunk_units_min_length"),
                        cfg.data.get("sample_units_max_length"))
    val = load_chunks(cfg.data.val_path, tok, ctx)

    from slamkit_b200.lm import B200UnitLM, LMConfig
    from slamkit_b200.trainer import B200Trainer
    if cfg.model.tlm_type not in ("twist", "gslm", "b200"):
        raise ValueError(f"Unknown slm type: {cfg.model.tlm_type}")                        # token_lm.py:43
    bs, ga = ta.per_device_train_batch_size, ta.gradient_accumulation_steps
    try:
        from slamkit_b200.integration import tlm_b200_from_cfg
        model = tlm_b200_from_cfg(to_container(cfg.model), device=f"cuda:{local_rank}", max_batch=bs)
    except Exception as e:   # offline: base model config / weights unreachable -> Qwen2.5-0.5B-shaped random init
        logger.warning(f"base model '{cfg.model.config_args.base_model_name}' unavailable ({type(e).__name__}); random init")
        lm_cfg = LMConfig(vocab_size=cfg.model.config_args.vocab_size, rope_theta=float(cfg.model.config_args.get("rope_theta", 10000)))
        for k in ("hidden", "n_layers", "n_heads", "n_kv_heads", "ffn"):
            if cfg.model.get("shape", {}).get(k) is not None:
                setattr(lm_cfg, k, cfg.model.shape[k])
        model = B200UnitLM(lm_cfg, device=f"cuda:{local_rank}", max_batch=bs, max_seq=ctx, seed=0)

    steps_per_epoch = max(1, math.ceil(len(train) / (bs * ga * world)))
    total_steps = ta.get("max_steps") or int(steps_per_epoch * ta.num_train_epochs)
    warmup = ta.get("warmup_steps", 0)
    if warmup > 0 and ta.get("warmup_ratio", 0.0) > 0 and total_steps * ta.warmup_ratio > warmup:
        warmup = int(math.ceil(total_steps * ta.warmup_ratio))                             # cli/train.py:48-54
    min_lr = (ta.get("lr_scheduler_kwargs") or {}).get("min_lr", 0.0)
    trainer = B200Trainer(model, lr=ta.learning_rate, min_lr=min_lr, warmup_steps=warmup, total_steps=total_steps,
                          max_grad_norm=ta.max_grad_norm, weight_decay=ta.get("weight_decay", 0.0), grad_accum=ga)
    budget = parse_run_time(cfg.run_time) if cfg.get("run_time") is not None else None
    max_tokens = cfg.get("train_max_tokens")
    g = torch.Generator().manual_seed(ta.get("seed", 42))
    order = torch.randperm(len(train), generator=g).tolist()
    t0, step, cursor = time.time(), 0, rank * bs
    log = []
    while step < total_steps:
        micro = []
        for _ in range(ga):
            idx = [order[(cursor + i) % len(order)] for i in range(bs)]
            cursor += bs * world
            micro.append(collate([train[i] for i in idx], tok.pad_token_id))
        loss = trainer.train_step(micro)
        step += 1
        if step % ta.get("logging_steps", 10) == 0 or step == total_steps:
            rec = {"step": step, "loss": float(loss), "tokens_seen": int(trainer.tokens_seen), "elapsed_s": time.time() - t0}
            log.append(rec)
            if rank == 0:
                print(json.dumps(rec), flush=True)
        if budget is not None and time.time() - t0 > budget:
            break
        if max_tokens is not None and int(trainer.tokens_seen) * world >= max_tokens:
            break
    # evaluation loss on the validation chunks (forward only)
    ev = []
    for i in range(0, len(val), ta.per_device_eval_batch_size):
        b = collate(val[i:i + ta.per_device_eval_batch_size], tok.pad_token_id)
        ev.append(float(model.forward(b["input_ids"], labels=b["labels"]).loss))
    if rank == 0:
        os.makedirs(ta.output_dir, exist_ok=True)
        model.save_pretrained(ta.output_dir, base_model_name=cfg.model.config_args.base_model_name)   # HF UnitLM layout
        tok.save_pretrained(ta.output_dir)
        json.dump({"log": log, "eval_loss": sum(ev) / max(1, len(ev)), "steps": step}, open(os.path.join(ta.output_dir, "trainer_state.json"), "w"))
        print(json.dumps({"eval_loss": sum(ev) / max(1, len(ev)), "steps": step}), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return log


if __name__ == "__main__":
    main()
