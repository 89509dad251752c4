"""Shared skeleton of the two trainers. The reference duplicates ~300 lines between
dalm/training/rag_e2e/train_rage2e.py:229-527 and dalm/training/retriever_only/train_retriever_only.py:175-421; here the
common part (tokenise -> DataLoader -> Adam + scheduler -> resume -> epoch/step loop -> checkpoints -> final artefacts)
is one function parameterised by a small `Recipe`, and the per-step work is dalm_b200's fused launch sequence.
"""
from __future__ import annotations

import argparse
import math
import os
import random
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import torch
from torch.utils.data import DataLoader

from ...accel import Accelerator, get_logger, set_seed
from ...optim import FusedAdam
from .train_utils import load_model_hook, save_model_hook

logger = get_logger(__name__)

# Optional measurement hook (bench.py --through-trainer): {"warmup": W, "steps": K} -> after the run also "seconds" (wall clock
# between device synchronisations at optimizer steps W and W+K) and "steps_timed". None = off: no synchronisation is added.
STEP_PROBE: Optional[Dict[str, Any]] = None


@dataclass
class Recipe:
    title: str                                   # log banner
    tracker_project: str                         # accelerator.init_trackers name (reference :368 / :306)
    build_model: Callable[[], torch.nn.Module]
    tokenize: Callable[[torch.nn.Module, Any], Any]            # (model, raw_dataset) -> tokenised dataset
    step: Callable[[torch.nn.Module, Dict[str, torch.Tensor], float, float], Dict[str, torch.Tensor]]
    banks: Callable[[torch.nn.Module], List[Any]]              # trainable LoRA banks
    repack: Callable[[torch.nn.Module], None]
    save_final: Callable[[torch.nn.Module, str], None]
    map_num_proc: Optional[int] = None
    step_fn: Any = None                          # the underlying fused_*_step (signature of train_utils.fused_rag_step)


def collate(features: List[Dict[str, Any]]) -> Dict[str, torch.Tensor]:
    """transformers.default_data_collator for this path's features: python int lists -> int64 tensors (reference :331)"""
    return {k: torch.tensor([f[k] for f in features], dtype=torch.int64) for k in features[0]}


def parse_resume(path: str, steps_per_epoch: int, loader_len: int, gas: int) -> Tuple[int, Optional[int], int]:
    """`epoch_{i}` / `step_{i}` directory names -> (starting_epoch, resume_step, completed_steps); reference :399-411"""
    tag = os.path.splitext(os.path.basename(os.path.normpath(path)))[0]
    if "epoch" in tag:
        start = int(tag.replace("epoch_", "")) + 1
        return start, None, start * steps_per_epoch
    resume = int(tag.replace("step_", "")) * gas
    start = resume // loader_len
    resume -= start * loader_len
    return start, resume, resume // gas


def plan_schedule(unsharded_loader_len: int, gas: int, num_train_epochs: int, max_train_steps: Optional[int]) -> Tuple[int, int]:
    """(max_train_steps, num_train_epochs) exactly as the reference derives them (train_rage2e.py:339-357) - from the loader
    length BEFORE accelerator.prepare shards it, so the epoch count does not depend on the number of ranks"""
    steps_per_epoch = math.ceil(unsharded_loader_len / gas)
    if max_train_steps is None:
        max_train_steps = num_train_epochs * steps_per_epoch
    return max_train_steps, math.ceil(max_train_steps / steps_per_epoch)


def run_training(recipe: Recipe, *, dataset_or_path: Any, per_device_train_batch_size: int, learning_rate: float,
                 logit_scale: float, num_train_epochs: int, max_train_steps: Optional[int],
                 gradient_accumulation_steps: int, lr_scheduler_type: Any, num_warmup_steps: int,
                 output_dir: Optional[str], seed: Optional[int], checkpointing_steps: Optional[Any],
                 resume_from_checkpoint: Optional[str], with_tracking: bool, report_to: str,
                 config_for_tracker: Dict[str, Any]) -> Dict[str, Any]:
    from transformers import get_scheduler
    from tqdm.auto import tqdm

    from ...utils import load_dataset

    accelerator = (Accelerator(log_with=report_to, project_dir=output_dir,
                               gradient_accumulation_steps=gradient_accumulation_steps)
                   if with_tracking else Accelerator(gradient_accumulation_steps=gradient_accumulation_steps))
    logger.info(accelerator.state, main_process_only=False)
    if seed is not None:
        set_seed(seed)
    if accelerator.is_main_process and output_dir is not None:
        os.makedirs(output_dir, exist_ok=True)
    accelerator.wait_for_everyone()

    model = recipe.build_model()
    dataset = load_dataset(dataset_or_path)
    processed = recipe.tokenize(model, dataset)
    for index in random.sample(range(len(processed)), min(2, len(processed))):
        logger.info(f"Sample {index} of the training set: {processed[index]}.")

    gen = torch.Generator()
    gen.manual_seed(seed if seed is not None else 0)          # same permutation on every rank (accelerate semantics)
    from . import packed_dataset
    if packed_dataset.enabled():
        # opt-in input pipeline: the tokenised set as one memory-mapped int32 matrix, a batch = one gather; same
        # sampler, same shards, the same tensors as `collate` builds (packed_dataset.py)
        cache_root = os.path.join(output_dir, ".packed_cache") if output_dir is not None else None
        packed = packed_dataset.PackedDataset.from_rows(processed, packed_dataset.PackedDataset.cache_path(processed, cache_root)
                                                        if accelerator.num_processes == 1 else None)
        loader = DataLoader(packed, shuffle=True, collate_fn=packed.collate, batch_size=per_device_train_batch_size,
                            pin_memory=torch.cuda.is_available(), generator=gen)
    else:
        loader = DataLoader(processed, shuffle=True, collate_fn=collate, batch_size=per_device_train_batch_size,
                            pin_memory=torch.cuda.is_available(), generator=gen)

    optimizer = FusedAdam(model.parameters(), lr=learning_rate)     # Adam, no weight decay (reference ignores the flag)
    max_train_steps, num_train_epochs = plan_schedule(len(loader), gradient_accumulation_steps, num_train_epochs, max_train_steps)
    sched_name = getattr(lr_scheduler_type, "value", lr_scheduler_type)
    scheduler = get_scheduler(name=sched_name, optimizer=optimizer, num_warmup_steps=num_warmup_steps,
                              num_training_steps=max_train_steps)
    # The reference does all of its step / epoch arithmetic BEFORE accelerator.prepare (train_rage2e.py:339-357 vs :416):
    # the epoch count comes from the UNSHARDED loader length, so `num_train_epochs=E` means E passes over the data on any
    # number of ranks (each pass len/W optimizer steps per rank, the wrapped scheduler stepping W times per update so the LR
    # schedule still ends with the last epoch); `max_train_steps` keeps its unsharded value as the early-stop bound.
    model, optimizer, loader, scheduler = accelerator.prepare(model, optimizer, loader, scheduler)
    steps_per_epoch = math.ceil(len(loader) / gradient_accumulation_steps)       # per rank: resume arithmetic, logging
    if checkpointing_steps is not None and str(checkpointing_steps).isdigit():
        checkpointing_steps = int(checkpointing_steps)
    if with_tracking:
        accelerator.init_trackers(recipe.tracker_project, config_for_tracker)
    accelerator.register_save_state_pre_hook(save_model_hook)
    accelerator.register_load_state_pre_hook(load_model_hook)

    total_bs = per_device_train_batch_size * accelerator.num_processes * gradient_accumulation_steps
    logger.info(f"***** {recipe.title} *****")
    logger.info(f"  Num examples = {len(processed)}")
    logger.info(f"  Num Epochs = {num_train_epochs}")
    logger.info(f"  Instantaneous batch size per device = {per_device_train_batch_size}")
    logger.info(f"  Total train batch size (w. parallel, distributed & accumulation) = {total_bs}")
    logger.info(f"  Gradient Accumulation steps = {gradient_accumulation_steps}")
    logger.info(f"  Total optimization steps = {max_train_steps}")

    progress = tqdm(range(max_train_steps), disable=not accelerator.is_local_main_process)
    completed, start_epoch, resume_step = 0, 0, None
    if resume_from_checkpoint:
        logger.info(f"Resumed from checkpoint: {resume_from_checkpoint}")
        accelerator.load_state(resume_from_checkpoint)
        recipe.repack(model)
        start_epoch, resume_step, completed = parse_resume(resume_from_checkpoint, steps_per_epoch, len(loader),
                                                           gradient_accumulation_steps)
    progress.update(completed)

    banks = recipe.banks(model)
    sync = accelerator.gradient_sync(banks)        # ONE collective per optimizer step: both banks' gradients + the loss scalar
    last_loss = None
    use_graph = os.environ.get("DALM_B200_CUDA_GRAPH", "1") != "0" and torch.cuda.is_available()
    if sync.overlaps_backward:          # full fine-tuning on > 1 rank: per-layer all-reduces are issued DURING the backward
        use_graph = False               # (DDP bucket semantics), which a single captured graph cannot contain
    from . import negatives
    if negatives.active():              # cross-rank negatives: an all-gather sits in the middle of the launch sequence
        use_graph = False
        logger.info("DALM_B200_CROSS_RANK_NEGATIVES=1: in-batch negatives are gathered over all ranks (not the reference's semantics)")
    graphed = None
    for epoch in range(start_epoch, num_train_epochs):
        model.train()
        total_loss = torch.zeros((), dtype=torch.float32, device=accelerator.device)
        active = loader
        if resume_from_checkpoint and epoch == start_epoch and resume_step is not None:
            active = accelerator.skip_first_batches(loader, resume_step)
        accelerator._loader = active                # `accumulate` reads end_of_dataloader from the loader being iterated
        for step, batch in enumerate(active):
            with accelerator.accumulate(model):
                sync.armed = accelerator.sync_gradients     # accumulation micro-steps keep their gradients local
                if use_graph and graphed is None:
                    from .train_utils import GraphedStep
                    try:
                        graphed = GraphedStep(recipe.step_fn, model, batch, float(logit_scale),
                                              1.0 / gradient_accumulation_steps, zero_grads=optimizer.zero_grad)
                    except Exception as e:                       # capture is an optimisation, never a requirement
                        logger.warning(f"CUDA-graph capture of the step failed ({type(e).__name__}: {e}); running eagerly")
                        use_graph = False
                out = (graphed(batch) if graphed is not None else
                       recipe.step(model, batch, float(logit_scale), 1.0 / gradient_accumulation_steps))
                if accelerator.sync_gradients:                  # gradients averaged + loss rank-SUMMED (reference :469) in one all-reduce
                    total_loss += sync.reduce(out["loss"])
                else:                                           # accumulating micro-step: only the logged loss crosses ranks
                    total_loss += accelerator.reduce(out["loss"].detach().float(), reduction="sum")
                if accelerator.sync_gradients:
                    optimizer.step()
                    recipe.repack(model)
                scheduler.step()
                if accelerator.sync_gradients:
                    optimizer.zero_grad()
            if accelerator.sync_gradients:
                progress.update(1)
                completed += 1
                if STEP_PROBE is not None and completed in (STEP_PROBE["warmup"], STEP_PROBE["warmup"] + STEP_PROBE["steps"]):
                    import time
                    torch.cuda.synchronize()
                    if completed == STEP_PROBE["warmup"]:
                        STEP_PROBE["t0"] = time.perf_counter()
                    else:
                        STEP_PROBE["seconds"] = time.perf_counter() - STEP_PROBE["t0"]
                        STEP_PROBE["steps_timed"] = STEP_PROBE["steps"]
            if (step + 1) % 100 == 0:
                last_loss = (total_loss / (step + 1)).item()
                logger.info(f"Step: {step + 1}, Loss: {last_loss}")
                if with_tracking:
                    accelerator.log({"train/loss": last_loss}, step=completed)
            if isinstance(checkpointing_steps, int) and checkpointing_steps > 0:
                if completed % checkpointing_steps == 0 and output_dir is not None and accelerator.sync_gradients:
                    accelerator.save_state(os.path.join(output_dir, f"step_{completed}"))
            if completed >= max_train_steps:
                break
        epoch_loss = total_loss.item() / max(len(loader), 1)
        last_loss = epoch_loss
        logger.info(f"epoch {epoch}: train/epoch_loss = {epoch_loss}")
        if with_tracking:
            accelerator.log({"train/epoch_loss": epoch_loss}, step=completed)
        if output_dir is not None:
            accelerator.wait_for_everyone()
            if isinstance(checkpointing_steps, str):
                accelerator.save_state(os.path.join(output_dir, f"epoch_{epoch}"))
            if accelerator.is_main_process:
                recipe.save_final(model, output_dir)
            accelerator.wait_for_everyone()
    if with_tracking:
        accelerator.end_training()
    return {"completed_steps": completed, "last_loss": last_loss, "model": model}


# ----------------------------------------------------------------------------------------------------------------
# argparse tables (the script entry points keep the reference's flag names and ITS argparse defaults, which differ from
# the function / CLI defaults in several places: SURVEY §5 "Config / flags")
# ----------------------------------------------------------------------------------------------------------------
def build_parser(description: str, flags: Sequence[Tuple[str, Dict[str, Any]]]) -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description=description)
    for name, kw in flags:
        parser.add_argument(f"--{name}", **kw)
    return parser
