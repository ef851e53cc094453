"""Coach — the reference's training driver (reference training/coach.py:36-834) on the B200 hot path, same constructor,
step semantics and method names:

    Coach(cfg).train()                                                                     scripts/train.py:23-24
    coach.py:155-156  mode 3: train_dataset.reset_sampled_object() before every micro-step (one object per batch)
    coach.py:158      accelerator.accumulate(...)  -> gradient accumulation window (optim.gradient_accumulation_steps)
    coach.py:165-169  latents = vae.encode(pixel_values).latent_dist.sample().detach() * scaling_factor   [models/vae.py]
    coach.py:172-183  noise, timesteps ~ U[0, T), noisy_latents = scheduler.add_noise(latents, noise, timesteps)
    coach.py:186-194  _hs = self.get_text_conditioning(...)            -> context dict (XTI protocol)   [batched CUDA path]
    coach.py:197-198  model_pred = self.unet(noisy_latents, timesteps, _hs).sample      [CUDA library]
    coach.py:201-209  target = noise | scheduler.get_velocity(...)
    coach.py:211-214  loss = F.mse_loss(model_pred.float(), target.float()); backward   [CUDA dgrad-only backward]
    coach.py:216-218  optimizer.step(); lr_scheduler.step(); optimizer.zero_grad()      (on the window's last micro-step)

`Coach(cfg)` builds tokenizer / scheduler / text encoder / VAE / UNet / dataset / mappers / optimizer from the RunConfig
exactly in the reference's order (coach.py:38-111).  Every piece can also be injected (`Coach(cfg, unet=..., conditioning=...)`),
which is how the tests and bench.py run steps on seeded synthetic weights.  Multi-GPU is batch-parallel: one process per GPU
and ONE all-reduce of the flat trainable-gradient buffer per optimiser step (training/dist.py) instead of accelerate's DDP
wrapper (coach.py:97-99); inside an accumulation window nothing is exchanged (DDP no_sync semantics).
"""
from __future__ import annotations

import itertools
import math
from pathlib import Path
from types import SimpleNamespace
from typing import Callable, Dict, Iterable, List, Optional, Tuple

import torch
import torch.nn.functional as F

from ..constants import UNET_LAYERS
from ..schedulers import DDPMScheduler
from .dist import FlatGradAllReducer


def _get(obj, dotted: str, default=None):
    for part in dotted.split("."):
        if obj is None:
            return default
        obj = obj.get(part, None) if isinstance(obj, dict) else getattr(obj, part, None)
    return default if obj is None else obj


def _randn(shape, generator, device, dtype=torch.float32):
    """diffusers' randn_tensor: draw on the generator's device (a CPU generator is legal) and move."""
    gdev = generator.device if generator is not None else device
    x = torch.randn(shape, generator=generator, device=gdev, dtype=dtype)
    return x.to(device)


class Coach:

    def __init__(self, cfg, unet=None, conditioning: Optional[Callable[..., Dict]] = None,
                 noise_scheduler: Optional[DDPMScheduler] = None, optimizer: Optional[torch.optim.Optimizer] = None,
                 lr_scheduler=None, generator: Optional[torch.Generator] = None, vae=None, tokenizer=None, text_encoder=None,
                 train_dataset=None, device="cuda"):
        self.cfg = cfg
        self.generator = generator
        self.global_step = 0          # optimiser steps taken
        self.micro_step = 0           # forward/backward passes taken
        self.device = torch.device(device)
        self.accumulation_steps = max(1, int(_get(cfg, "optim.gradient_accumulation_steps", 1) or 1))
        self.sync_gradients = True
        self.learnable_mode = int(_get(cfg, "learnable_mode", 2))
        self.train_dataset = train_dataset
        self.train_dataloader = None
        if conditioning is not None or unet is not None:
            # ---- injected components (tests, bench.py, scripts/*): the caller built the models ----
            assert unet is not None and conditioning is not None, "give both unet and conditioning, or neither"
            self.unet, self.vae, self.tokenizer = unet, vae, tokenizer
            self.text_encoder = text_encoder
            self.conditioning = conditioning
            self.noise_scheduler = noise_scheduler or DDPMScheduler()
        else:
            # ---- the reference's constructor, step by step (coach.py:38-111) ----
            seed = _get(cfg, "optim.seed")
            if seed is not None:
                torch.manual_seed(int(seed))
            self.tokenizer, self.noise_scheduler, self.text_encoder, self.vae, self.unet = self._init_sd_models()
            if noise_scheduler is not None:
                self.noise_scheduler = noise_scheduler
            self.train_dataset = train_dataset or self._init_dataset()
            self.train_dataloader = self._init_dataloader(self.train_dataset)
            ds = self.train_dataset
            self.placeholder_object_tokens, self.placeholder_view_tokens = ds.placeholder_object_tokens, ds.placeholder_view_tokens
            self.placeholder_tokens, self.fixed_object_token = ds.placeholder_tokens, ds.fixed_object_token
            (self.token_embeds, self.placeholder_token_ids, self.placeholder_view_token_ids,
             self.placeholder_object_token_ids) = Coach._add_concept_token_to_tokenizer_static(
                cfg, ds.placeholder_view_tokens, ds.placeholder_object_tokens, self.tokenizer, self.text_encoder)
            cfg.data.placeholder_view_tokens = self.placeholder_view_tokens
            lookup, mapper_view, self.loaded_iteration = self._init_neti_mapper()
            self.text_encoder.text_model.embeddings.set_mapper(lookup, mapper_view)
            self.conditioning = self.text_encoder.conditioning
            self._freeze_all_modules()
            self._set_attn_processor()
        params = self._trainable_params()
        self.optimizer = optimizer if optimizer is not None else (self._init_optimizer(params) if params else None)
        self.lr_scheduler = lr_scheduler if lr_scheduler is not None else \
            (self._init_scheduler(self.optimizer) if self.optimizer is not None and conditioning is None else None)
        self.reducer = FlatGradAllReducer(params) if params else None
        if self.reducer is not None:
            self.reducer.broadcast_parameters_(0)       # what accelerate's DDP wrapper does at coach.py:97-99
        if conditioning is None:
            from ..checkpoint_handler import CheckpointHandler
            from .config import to_dict
            self.checkpoint_handler = CheckpointHandler(
                cfg=to_dict(cfg), placeholder_view_tokens=self.placeholder_view_tokens,
                placeholder_view_token_ids=self.placeholder_view_token_ids,
                placeholder_object_tokens=self.placeholder_object_tokens,
                placeholder_object_token_ids=self.placeholder_object_token_ids, save_root=_get(cfg, "log.exp_dir", "."))

    # ---- construction (reference names) ---------------------------------------------------------------------------------
    def _init_sd_models(self):
        """coach.py:600-640.  `model.pretrained_model_name_or_path: synthetic` -> seeded weights at the SD-2.1 shapes (no
        checkpoint exists offline); a local diffusers directory -> its unet / vae (text encoder + tokenizer files must be
        there too; nothing is ever downloaded)."""
        name = str(_get(self.cfg, "model.pretrained_model_name_or_path", "synthetic"))
        if name.startswith("synthetic"):
            from .synthetic import build_sd_models
            from ..sd21 import SD21, TINY
            from ..models.clip_encoder import SD21_TEXT, ClipEncoderConfig
            tiny = name.endswith("-tiny")
            text_cfg = ClipEncoderConfig(hidden_size=TINY.cross_attention_dim, num_attention_heads=2, num_hidden_layers=2,
                                         intermediate_size=256) if tiny else SD21_TEXT
            return build_sd_models(self.device, seed=int(_get(self.cfg, "seed", 0) or 0), unet_cfg=TINY if tiny else SD21,
                                   text_cfg=text_cfg, with_vae=True)
        root = Path(name)
        if not root.is_dir():
            raise FileNotFoundError(f"model.pretrained_model_name_or_path='{name}' is not a local directory and this build never "
                                    f"downloads; use 'synthetic' for seeded weights at the SD-2.1 shapes")
        from ..models.vae import AutoencoderKL
        from ..unet import UNet2DConditionModel
        from transformers import CLIPTokenizer
        tokenizer = CLIPTokenizer.from_pretrained(str(root), subfolder="tokenizer")
        unet = UNet2DConditionModel.from_pretrained(str(root), subfolder="unet", device=self.device)
        vae = AutoencoderKL.from_pretrained(str(root), subfolder="vae", device=self.device)
        text_encoder = self._load_text_encoder(root / "text_encoder")
        return tokenizer, DDPMScheduler(_get(self.cfg, "model.prediction_type", "v_prediction")), text_encoder, vae, unet

    def _load_text_encoder(self, root: Path):
        """transformers CLIPTextModel weights (text_model.{embeddings,encoder,final_layer_norm}.*) -> NeTICLIPTextModel."""
        from ..models.clip_encoder import SD21_TEXT, CLIPEncoder
        from ..models.neti_clip_text_encoder import NeTICLIPTextModel
        sd = None
        for fn in ("pytorch_model.bin", "model.safetensors"):
            p = root / fn
            if p.exists():
                if fn.endswith(".bin"):
                    sd = torch.load(p, map_location="cpu")
                else:
                    from safetensors.torch import load_file
                    sd = load_file(str(p))
                break
        if sd is None:
            raise FileNotFoundError(f"no text encoder weights under {root}")
        pre = "text_model.encoder."
        enc = CLIPEncoder({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}, SD21_TEXT, self.device)
        return NeTICLIPTextModel.from_parts(sd["text_model.embeddings.token_embedding.weight"],
                                            sd["text_model.embeddings.position_embedding.weight"],
                                            (sd["text_model.final_layer_norm.weight"], sd["text_model.final_layer_norm.bias"]), enc)

    def _init_dataset(self):
        """coach.py:682-702.  The DTU / mode-0 folder readers (training/dataset.py) are CPU data preparation outside the hot
        path; `data.train_data_dir: synthetic` yields items of the same format."""
        d = self.cfg.data
        if str(d.train_data_dir) != "synthetic":
            raise NotImplementedError("only data.train_data_dir='synthetic' is built in; pass train_dataset=<a dataset yielding the "
                                      "reference's item dict> for real data (training/dataset.py:605-739)")
        from .synthetic import SyntheticTIDataset
        size = (int(d.resolution), int(d.resolution))
        return SyntheticTIDataset(self.learnable_mode, self.tokenizer, placeholder_object_token=d.placeholder_object_token,
                                  placeholder_object_tokens=d.placeholder_object_tokens, train_data_subsets=d.train_data_subsets,
                                  fixed_object_token=d.fixed_object_token_or_path or "object",
                                  camera_representation=d.camera_representation, size=size,
                                  length=max(8, int(d.repeats)), seed=int(_get(self.cfg, "seed", 0) or 0))

    def _init_dataloader(self, dataset):
        return torch.utils.data.DataLoader(dataset, batch_size=int(self.cfg.optim.train_batch_size), shuffle=True,
                                           num_workers=0)

    @staticmethod
    def _add_concept_token_to_tokenizer_static(cfg, placeholder_view_tokens, placeholder_object_tokens, tokenizer, text_encoder):
        """coach.py:320-397: registers the placeholder tokens, initialises their embedding rows from the super-category words
        and records the target norms in cfg.model.  Modifies tokenizer, text_encoder and cfg in place."""
        placeholder_tokens = list(placeholder_view_tokens) + list(placeholder_object_tokens)
        if tokenizer.add_tokens(placeholder_tokens) == 0:
            raise ValueError("No new tokens were added to the tokenizer. Please pass a different `placeholder_token` that is "
                             "not already in the tokenizer.")
        view_ids = tokenizer.convert_tokens_to_ids(list(placeholder_view_tokens))
        object_ids = tokenizer.convert_tokens_to_ids(list(placeholder_object_tokens))
        all_ids = tokenizer.convert_tokens_to_ids(placeholder_tokens)
        assert set(view_ids) | set(object_ids) == set(all_ids)
        sup_obj = tokenizer.encode(cfg.data.super_category_object_token, add_special_tokens=False)
        sup_view = tokenizer.encode(cfg.data.super_category_view_token, add_special_tokens=False)
        if len(sup_obj) != 1:
            raise ValueError(f"object supercategory [{cfg.data.super_category_object_token}] not in the vocabulary")
        if len(sup_view) != 1:
            raise ValueError(f"view supercategory [{cfg.data.super_category_view_token}] not in the vocabulary")
        sup_obj, sup_view = sup_obj[0], sup_view[0]
        text_encoder.resize_token_embeddings(len(tokenizer))
        token_embeds = text_encoder.get_input_embeddings().weight.data
        if view_ids:
            token_embeds[view_ids] = token_embeds[sup_view].clone().unsqueeze(0).repeat(len(view_ids), 1)
        if object_ids:
            token_embeds[object_ids] = token_embeds[sup_obj].clone().unsqueeze(0).repeat(len(object_ids), 1)
        cfg.model.target_norm_view = cfg.model.target_norm_object = None
        if cfg.model.normalize_view_mapper_output:
            if sup_view == tokenizer.unk_token_id:
                raise ValueError(f"super_category_view_token [{cfg.data.super_category_view_token}] is unknown to the tokenizer")
            cfg.model.target_norm_view = token_embeds[sup_view].norm().item()
        if cfg.model.normalize_object_mapper_output:
            if sup_obj == tokenizer.unk_token_id:
                raise ValueError(f"super_category_object_token [{cfg.data.super_category_object_token}] is unknown to the tokenizer")
            cfg.model.target_norm_object = token_embeds[sup_obj].norm().item()
        return token_embeds, all_ids, view_ids, object_ids

    def _init_neti_mapper(self):
        """coach.py:492-598: M_o for modes 0,2,3,4,5 (one per object token); fresh M_v for 1,2,3; loaded M_v for 4,5."""
        from ..models.neti_mapper import NeTIMapper
        cfg, m = self.cfg, self.cfg.model
        if self.learnable_mode not in (0, 1, 2, 3, 4, 5):
            raise NotImplementedError()
        if m.mapper_checkpoint_path:
            raise NotImplementedError("Check this implementation is right")          # coach.py:500-501: resume is not implemented
        norm = lambda v: None if v is None else torch.tensor(float(v))     # noqa: E731
        common = dict(output_dim=m.word_embedding_dim, arch_mlp_hidden_dims=m.arch_mlp_hidden_dims,
                      use_nested_dropout=m.use_nested_dropout, nested_dropout_prob=m.nested_dropout_prob,
                      num_pe_time_anchors=m.num_pe_time_anchors, pe_sigmas=m.pe_sigmas, arch_view_net=m.arch_view_net,
                      arch_view_mix_streams=m.arch_view_mix_streams, arch_view_disable_tl=m.arch_view_disable_tl,
                      original_ti=m.original_ti)
        lookup, mapper_view = None, None
        if self.learnable_mode in (0, 2, 3, 4, 5):
            lookup = {}
            for tok, tid in zip(self.placeholder_object_tokens, self.placeholder_object_token_ids):
                lookup[tid] = NeTIMapper(embedding_type="object", placeholder_object_token=tok, norm_scale=norm(m.target_norm_object),
                                         use_positional_encoding=m.use_positional_encoding_object,
                                         output_bypass=m.output_bypass_object, output_bypass_alpha=m.output_bypass_alpha_object,
                                         bypass_unconstrained=m.bypass_unconstrained_object, **common)
        ds = self.train_dataset
        view_kw = dict(embedding_type="view", placeholder_view_tokens=self.placeholder_view_tokens,
                       placeholder_view_token_ids=self.placeholder_view_token_ids, norm_scale=norm(m.target_norm_view),
                       use_positional_encoding=m.use_positional_encoding_view, output_bypass=m.output_bypass_view,
                       output_bypass_alpha=m.output_bypass_alpha_view, bypass_unconstrained=m.bypass_unconstrained_view,
                       cam_mins=getattr(ds, "cam_mins", None), cam_maxs=getattr(ds, "cam_maxs", None), **common)
        if self.learnable_mode in (1, 2, 3):
            mapper_view = NeTIMapper(**view_kw)
        elif self.learnable_mode in (4, 5):
            from ..checkpoint_handler import CheckpointHandler
            _, mapper_view = CheckpointHandler.load_mapper(m.pretrained_view_mapper, "view",
                                                           placeholder_view_tokens=self.placeholder_view_tokens,
                                                           placeholder_view_token_ids=self.placeholder_view_token_ids,
                                                           cam_mins=getattr(ds, "cam_mins", None),
                                                           cam_maxs=getattr(ds, "cam_maxs", None))
        return lookup, mapper_view, None

    def _freeze_all_modules(self):
        """coach.py:642-677.  UNet / VAE / text model are frozen by construction on this path; the mappers train.  Mode 5
        keeps the pretrained M_v out of the optimiser; here it is also frozen for real (`requires_grad_(False)`), so its
        gradients are not computed at all - the reference computes and discards them (coach.py:661-669)."""
        if self.vae is not None:
            self.vae.requires_grad_(False)
        self.unet.requires_grad_(False)
        tm = self.text_encoder.text_model
        tm.encoder.requires_grad_(False)
        tm.final_layer_norm.requires_grad_(False)
        tm.embeddings.position_embedding.requires_grad_(False)
        for mapper in tm.embeddings.mapper_object_lookup.values():
            mapper.requires_grad_(self.learnable_mode in (0, 2, 3, 4, 5))
            mapper.train()
        if tm.embeddings.mapper_view is not None:
            tm.embeddings.mapper_view.requires_grad_(self.learnable_mode in (1, 2, 3, 4))
            tm.embeddings.mapper_view.train(self.learnable_mode in (1, 2, 3, 4))
        if _get(self.cfg, "optim.gradient_checkpointing", False):
            self.text_encoder.gradient_checkpointing_enable()
            self.unet.enable_gradient_checkpointing()

    def _set_attn_processor(self):
        from ..models.xti_attention_processor import XTIAttenProc
        self.unet.set_attn_processor(XTIAttenProc())

    def _trainable_params(self) -> List[torch.nn.Parameter]:
        """coach.py:735-748: object mappers in every mode but 1, the view mapper in modes 1-4."""
        c = self.conditioning
        if not isinstance(c, torch.nn.Module):
            return []
        return [p for p in c.parameters() if p.requires_grad]

    def _world(self) -> int:
        return self.reducer.world if getattr(self, "reducer", None) is not None else \
            (torch.distributed.get_world_size() if torch.distributed.is_available() and torch.distributed.is_initialized() else 1)

    def _init_optimizer(self, params) -> torch.optim.Optimizer:
        """coach.py:727-757: AdamW on the mapper parameters only, hyper-parameters from cfg.optim, lr scaled by accumulation x
        micro-batch x processes when optim.scale_lr."""
        o = _get(self.cfg, "optim") or self.cfg
        lr = float(_get(o, "learning_rate", 1e-3))
        if _get(o, "scale_lr", False):
            world = torch.distributed.get_world_size() if torch.distributed.is_available() and torch.distributed.is_initialized() else 1
            lr = lr * self.accumulation_steps * int(_get(o, "train_batch_size", 1)) * world
            try:
                o.learning_rate = lr                       # the reference writes the scaled value back into the config
            except AttributeError:
                pass
        return torch.optim.AdamW(params, lr=lr, betas=(float(_get(o, "adam_beta1", 0.9)), float(_get(o, "adam_beta2", 0.999))),
                                 weight_decay=float(_get(o, "adam_weight_decay", 1e-2)), eps=float(_get(o, "adam_epsilon", 1e-8)))

    def _init_scheduler(self, optimizer):
        """coach.py:759-770 (diffusers get_scheduler): warm-up and horizon are counted in micro-steps there because
        accelerate's wrapped scheduler is stepped every micro-step of a window; here it steps once per optimiser step, so
        both are counted in optimiser steps."""
        o = _get(self.cfg, "optim") or self.cfg
        kind = str(_get(o, "lr_scheduler", "constant"))
        warm, total = int(_get(o, "lr_warmup_steps", 0)), int(_get(o, "max_train_steps", 1000) or 1000)

        def lam(step: int) -> float:
            w = min(1.0, (step + 1) / warm) if warm > 0 and kind != "constant" else 1.0
            if kind in ("constant", "constant_with_warmup"):
                return w
            prog = min(1.0, max(0.0, (step - warm) / max(1, total - warm)))
            if kind == "linear":
                return w * (1.0 - prog)
            if kind in ("cosine", "cosine_with_restarts"):
                return w * 0.5 * (1.0 + math.cos(math.pi * prog))
            raise ValueError(f"unsupported optim.lr_scheduler '{kind}'")
        return torch.optim.lr_scheduler.LambdaLR(optimizer, lam)

    # ---- conditioning (same name / argument meaning as reference coach.py:276-283) ----------------------------------------
    def get_text_conditioning(self, input_ids=None, timesteps=None, input_ids_placeholder_object=None,
                              input_ids_placeholder_view=None, device=None, original_ti: bool = False) -> Dict:
        return self.conditioning(input_ids=input_ids, timesteps=timesteps,
                                 input_ids_placeholder_object=input_ids_placeholder_object,
                                 input_ids_placeholder_view=input_ids_placeholder_view, device=device, original_ti=original_ti)

    def encode_images(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """coach.py:165-169: images in [-1, 1] -> scaled latents, frozen VAE, no autograd history."""
        if self.vae is None:
            raise ValueError("Coach: batch carries pixel_values but no vae was given")
        dist = self.vae.encode(pixel_values.to(self.unet.device)).latent_dist
        return dist.sample(self.generator).detach() * self.vae.config.scaling_factor

    # ---- mode 3: one object per batch, the same on every rank -------------------------------------------------------------
    def reset_sampled_object(self) -> Optional[int]:
        """coach.py:155-156 -> dataset.py:584-600.  With several ranks, rank 0 draws and broadcasts the index, so that ONE
        object mapper is active in the step everywhere and the flat gradient buffer (M_v + that M_o) has the same layout on
        all ranks."""
        ds = self.train_dataset
        if ds is None or self.learnable_mode != 3:
            return None
        idx = ds.reset_sampled_object()
        if self._world() > 1:
            import torch.distributed as dist
            dev = self.unet.device if dist.get_backend() == "nccl" else "cpu"
            t = torch.tensor([idx], dtype=torch.int64, device=dev)
            dist.broadcast(t, src=0)
            idx = ds.reset_sampled_object(int(t))
        return idx

    def _active_params(self, batch: Dict) -> Optional[List[torch.nn.Parameter]]:
        """Parameters this step's prompts use: M_v (if trainable) + the object mapper named by the batch."""
        c = self.conditioning
        lookup = getattr(c, "mapper_object_lookup", None)
        if lookup is None or len(lookup) <= 1:
            return None                                   # a single object mapper: every trainable parameter is in use
        ph = batch.get("input_ids_placeholder_object")
        if ph is None:
            return None
        tid = int(ph[0]) if not torch.is_tensor(ph) else int(ph.reshape(-1)[0])
        act: List[torch.nn.Parameter] = []
        if getattr(c, "mapper_view", None) is not None:
            act += [p for p in c.mapper_view.parameters() if p.requires_grad]
        if str(tid) in lookup:
            act += [p for p in lookup[str(tid)].parameters() if p.requires_grad]
        return act

    # ---- one micro-step -------------------------------------------------------------------------------------------------
    def train_step(self, latents: Optional[torch.Tensor] = None, batch: Optional[Dict] = None) -> torch.Tensor:
        """One forward/backward pass; on the last pass of an accumulation window also all-reduce + optimizer step.
        `latents` given: the step starts at coach.py:172 (pre-encoded data); otherwise `batch["pixel_values"]` goes through
        the VAE first, as the reference does every step."""
        batch = batch or {}
        if latents is None:
            latents = self.encode_images(batch["pixel_values"])
        dev = latents.device
        noise = _randn(latents.shape, self.generator, dev, latents.dtype)
        bsz = latents.shape[0]
        gdev = self.generator.device if self.generator is not None else dev
        timesteps = torch.randint(0, self.noise_scheduler.config.num_train_timesteps, (bsz,), generator=self.generator,
                                  device=gdev).long().to(dev)
        noisy_latents = self.noise_scheduler.add_noise(latents, noise, timesteps)
        _hs = self.get_text_conditioning(input_ids=batch.get("input_ids"), timesteps=timesteps,
                                         input_ids_placeholder_object=batch.get("input_ids_placeholder_object"),
                                         input_ids_placeholder_view=batch.get("input_ids_placeholder_view"), device=dev,
                                         original_ti=bool(_get(self.cfg, "model.original_ti", False)))
        model_pred = self.unet(noisy_latents, timesteps, _hs).sample
        if self.noise_scheduler.config.prediction_type == "epsilon":
            target = noise
        elif self.noise_scheduler.config.prediction_type == "v_prediction":
            target = self.noise_scheduler.get_velocity(latents, noise, timesteps)
        else:
            raise ValueError(f"Unknown prediction type {self.noise_scheduler.config.prediction_type}")
        loss = F.mse_loss(model_pred.float(), target.float(), reduction="mean")
        # accelerator.backward divides by the window length so that the accumulated gradient is the window's mean
        (loss / self.accumulation_steps if self.accumulation_steps > 1 else loss).backward()
        self.micro_step += 1
        self.sync_gradients = self.micro_step % self.accumulation_steps == 0
        if self.sync_gradients:
            if self.reducer is not None:
                self.reducer.allreduce_(self._active_params(batch) if self.accumulation_steps == 1 else None)
            if self.optimizer is not None:
                self.optimizer.step()
                if self.lr_scheduler is not None:
                    self.lr_scheduler.step()
                self.optimizer.zero_grad(set_to_none=True)
            self.global_step += 1
        return loss.detach()

    # ---- the loop (coach.py:137-273) --------------------------------------------------------------------------------------
    def train(self, batches: Optional[Iterable] = None, max_train_steps: Optional[int] = None):
        """`Coach(cfg).train()` iterates the dataloader until optim.max_train_steps optimiser steps; an explicit iterable of
        latent tensors or of dicts in the dataloader's format (`pixel_values`, `input_ids`, ...) can be given instead."""
        max_steps = max_train_steps or _get(self.cfg, "optim.max_train_steps")
        losses = []
        save_steps = _get(self.cfg, "log.save_steps")

        def one(b):
            losses.append(self.train_step(batch=b) if isinstance(b, dict) else self.train_step(b))
            if self.sync_gradients and save_steps and getattr(self, "checkpoint_handler", None) is not None \
                    and self.global_step % int(save_steps) == 0 and self._is_main():
                self.save(f"learned_embeds-steps-{self.global_step}.bin", f"mapper-steps-{self.global_step}.pt")

        if batches is not None:
            for b in batches:
                one(b)
                if max_steps is not None and self.global_step >= max_steps:
                    break
            return losses
        assert self.train_dataloader is not None, "Coach was built from injected components: pass the batches to train()"
        while max_steps is None or self.global_step < max_steps:
            it = iter(self.train_dataloader)
            while True:
                if self.learnable_mode == 3:
                    self.reset_sampled_object()           # BEFORE the batch is drawn: its items name the sampled object
                try:
                    b = next(it)
                except StopIteration:
                    break
                losses.append(self.train_step(batch=b))
                if self.sync_gradients and save_steps and self.global_step % int(save_steps) == 0 and self._is_main():
                    self.save(f"learned_embeds-steps-{self.global_step}.bin", f"mapper-steps-{self.global_step}.pt")
                if max_steps is not None and self.global_step >= max_steps:
                    break
            if max_steps is None:
                break
        if self._is_main() and getattr(self, "checkpoint_handler", None) is not None and _get(self.cfg, "log.exp_dir") is not None:
            self.save("learned_embeds-final.bin", "mapper-final.pt")                  # coach.py:266-273
        return losses

    def _is_main(self) -> bool:
        d = torch.distributed
        return not (d.is_available() and d.is_initialized()) or d.get_rank() == 0

    def save(self, embeds_save_name: str, mapper_save_name: str) -> None:
        root = Path(self.checkpoint_handler.save_root)
        root.mkdir(parents=True, exist_ok=True)
        self.checkpoint_handler.save_model(self.conditioning, embeds_save_name, mapper_save_name)


class SyntheticConditioning(torch.nn.Module):
    """Stand-in for the NeTI mapper + CLIP path: a trainable table that emits the XTI context dict
    {"this_idx", "CONTEXT_TENSOR_i", "CONTEXT_TENSOR_BYPASS_i"} (one [B,77,D] pair per UNet layer).  Used by tests and
    smoke so that a step has real trainable parameters downstream of d_ctx without the text encoder."""

    def __init__(self, n_layers: int = len(UNET_LAYERS), context_len: int = 77, dim: int = 1024, rank: int = 8, seed: int = 0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.n_layers = n_layers
        self.base = torch.nn.Parameter(torch.randn(2, n_layers, context_len, rank, generator=g))
        self.proj = torch.nn.Parameter(torch.randn(rank, dim, generator=g) / rank ** 0.5)

    def forward(self, timesteps=None, device=None, **_) -> Dict:
        bsz = 1 if timesteps is None else timesteps.shape[0]
        ctx = (self.base @ self.proj)                                # [2, L, 77, D]
        out: Dict = {"this_idx": 0}
        for i in range(self.n_layers):
            out[f"CONTEXT_TENSOR_{i}"] = ctx[0, i].unsqueeze(0).expand(bsz, -1, -1)
            out[f"CONTEXT_TENSOR_BYPASS_{i}"] = ctx[1, i].unsqueeze(0).expand(bsz, -1, -1)
        return out
