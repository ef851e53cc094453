"""Seeded synthetic instance of the COMPLETE conditioning stack at the SD-2.1 shapes (no tokenizer / checkpoints exist
offline): token + position embeddings, the 23-layer CLIP encoder, one object mapper and one (theta, phi) view mapper
(141 696 parameters each), wired as `NeTIConditioning` - what `Coach(cfg, unet, conditioning=...)` needs to
run whole train steps.  Used by bench.py (`full_step`) and scripts/full_step_bench.py."""
from __future__ import annotations

from typing import Dict

import torch

from ..models.clip_encoder import SD21_TEXT, CLIPEncoder, ClipEncoderConfig, init_state_dict
from ..models.neti_conditioning import NeTIConditioning
from ..models.neti_mapper import NeTIMapper
from ..utils.types import PESigmas

OBJECT_TOKEN_ID = 49408
VIEW_TOKEN_IDS = [49409, 49410, 49411, 49412]
VIEW_TOKENS = ["<view_0_10_1p2>", "<view_10_40_1p2>", "<view_20_70_1p2>", "<view_35_100_1p2>"]


def build_conditioning(device="cuda", cfg: ClipEncoderConfig = SD21_TEXT, seed: int = 0, n_objects: int = 1) -> NeTIConditioning:
    """n_objects > 1: BASELINE config 4 (mode 3 multi-scene pretraining, train_m3.yaml): that many object mappers resident
    under consecutive placeholder ids OBJECT_TOKEN_ID + 16 + i (one is active per step), one shared view mapper."""
    g = torch.Generator().manual_seed(seed)
    C = cfg.hidden_size
    tok = torch.randn(49408 + 8 + (16 + n_objects if n_objects > 1 else 0), C, generator=g) * 0.02
    pos = torch.randn(77, C, generator=g) * 0.01
    enc = CLIPEncoder(init_state_dict(cfg, seed), cfg, device)
    sig = PESigmas(sigma_t=0.03, sigma_l=2.0, sigma_theta=0.5, sigma_phi=0.5, sigma_r=0.5, sigma_dtu12=0.5)
    kw = dict(output_dim=C, arch_mlp_hidden_dims=64, arch_view_net=15, arch_view_disable_tl=False, use_nested_dropout=False,
              pe_sigmas=sig, output_bypass=True, bypass_unconstrained=True, output_bypass_alpha=0.2)
    with torch.random.fork_rng(devices=[]):          # nn.Linear init draws from the process-local default generator
        torch.manual_seed(seed + 1)
        mo = NeTIMapper(embedding_type="object", norm_scale=torch.tensor(0.3714), placeholder_object_token="<statue>", **kw).to(device)
        mv = NeTIMapper(embedding_type="view", norm_scale=torch.tensor(0.4102), placeholder_view_tokens=list(VIEW_TOKENS),
                        placeholder_view_token_ids=list(VIEW_TOKEN_IDS), **kw).to(device)
        lookup = {OBJECT_TOKEN_ID: mo}
        if n_objects > 1:
            lookup = {object_token_id(i): NeTIMapper(embedding_type="object", norm_scale=torch.tensor(0.3714),
                                                     placeholder_object_token=f"<scan{i}>", **kw).to(device) for i in range(n_objects)}
    return NeTIConditioning(tok, pos, (torch.ones(C), torch.zeros(C)), enc, lookup, mv)


def object_token_id(i: int) -> int:
    """Placeholder id of object i of a multi-object (mode 3) synthetic conditioning stack."""
    return OBJECT_TOKEN_ID + 16 + i


def synthetic_prompt(batch: int, device="cuda", seed: int = 0, object_id: int = OBJECT_TOKEN_ID) -> Dict[str, torch.Tensor]:
    """`batch` prompts of 77 token ids holding the object placeholder and one view placeholder each."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1000, 40000, (batch, 77), generator=g)
    ids[:, 0], ids[:, 5] = 49406, object_id
    view = torch.tensor([VIEW_TOKEN_IDS[i % len(VIEW_TOKEN_IDS)] for i in range(batch)])
    ids[:, 3] = view
    # the placeholder ids are host metadata, as the reference's dataloader yields them (CPU tensors): the conditioning path
    # reads them as Python ints, and reading a CUDA tensor back would stall the host on the device every step
    return {"input_ids": ids.to(device), "input_ids_placeholder_object": torch.full((batch,), object_id),
            "input_ids_placeholder_view": view}


# ---------------------------------------------------------------------------------------------------------------------
# Offline stand-ins for what `Coach(cfg)` loads from the HF hub / from DTU folders in the reference (coach.py:600-640,
# 682-702).  No tokenizer vocabulary, checkpoint or dataset exists in this environment, so `model.pretrained_model_name_or_path:
# synthetic` / `data.train_data_dir: synthetic` select these.  They expose exactly the members the training path reads.
# ---------------------------------------------------------------------------------------------------------------------
class SyntheticTokenizer:
    """The CLIPTokenizer members the reference touches (coach.py:326-352, dataset.py:605-739, prompt_manager.py,
    sd_pipeline_call.py:128-140): add_tokens, convert_tokens_to_ids, encode, __call__(..., padding='max_length'),
    __len__, unk_token_id, model_max_length.  Words are whitespace-split; unseen words map to stable pseudo-ids."""
    model_max_length = 77
    bos_token_id, eos_token_id, unk_token_id = 49406, 49407, 49407
    pad_token_id = 49407

    def __init__(self, vocab_size: int = 49408):
        self._base = vocab_size
        self._added: Dict[str, int] = {}

    def __len__(self) -> int:
        return self._base + len(self._added)

    def add_tokens(self, tokens) -> int:
        n = 0
        for t in ([tokens] if isinstance(tokens, str) else tokens):
            if t not in self._added:
                self._added[t] = self._base + len(self._added)
                n += 1
        return n

    def _word_id(self, w: str) -> int:
        if w in self._added:
            return self._added[w]
        import zlib
        return 1000 + zlib.crc32(w.lower().encode()) % 39000          # stable, below the special tokens

    def convert_tokens_to_ids(self, tokens):
        if isinstance(tokens, str):
            return self._word_id(tokens)
        return [self._word_id(t) for t in tokens]

    def encode(self, text: str, add_special_tokens: bool = True):
        ids = [self._word_id(w) for w in text.split()]
        return [self.bos_token_id] + ids + [self.eos_token_id] if add_special_tokens else ids

    def __call__(self, text, padding="max_length", truncation=True, max_length=None, return_tensors="pt", **_):
        from types import SimpleNamespace
        L = max_length or self.model_max_length
        rows = []
        for t in ([text] if isinstance(text, str) else text):
            ids = self.encode(t)[:L]
            ids[-1] = self.eos_token_id if len(ids) == L else ids[-1]
            rows.append(ids + [self.pad_token_id] * (L - len(ids)))
        return SimpleNamespace(input_ids=torch.tensor(rows, dtype=torch.long))


class SyntheticTIDataset(torch.utils.data.Dataset):
    """Items in the reference dataset's format (training/dataset.py:605-739): pixel_values [3,H,W] in [-1,1], input_ids [77],
    input_ids_placeholder_object / _view (token ids, -1 without a view token), text, image_idx.  Modes as
    training/dataset.py:39-120: 0 object only; 1 view only (fixed object word); 2/4/5 view + one object; 3 view + several
    objects with ONE object per batch, redrawn by `reset_sampled_object()` (:584-600)."""

    def __init__(self, learnable_mode: int, tokenizer, placeholder_object_token: str = "<object>",
                 placeholder_object_tokens=None, train_data_subsets=None, fixed_object_token: str = "object",
                 camera_representation: str = "spherical", n_views: int = 6, size=(512, 512), length: int = 64, seed: int = 0):
        self.learnable_mode, self.tokenizer = learnable_mode, tokenizer
        self.size, self._length = size, length
        self.camera_representation = camera_representation
        g = torch.Generator().manual_seed(seed)
        if learnable_mode == 3:
            self.placeholder_object_tokens = list(placeholder_object_tokens or [f"<object{i}>" for i in range(14)])
            self.train_data_subsets = list(train_data_subsets or [f"scan{i}" for i in range(len(self.placeholder_object_tokens))])
        elif learnable_mode == 1:
            self.placeholder_object_tokens = []
            self.train_data_subsets = None
        else:
            self.placeholder_object_tokens = [placeholder_object_token]
            self.train_data_subsets = None
        self.fixed_object_token = fixed_object_token if learnable_mode == 1 else None
        self.cam_mins = self.cam_maxs = None
        if learnable_mode == 0:
            self.placeholder_view_tokens = []
        elif camera_representation == "dtu-12d":
            cams = torch.randn(n_views, 12, generator=g) * torch.tensor([1.0, 1, 1, 300] * 3) + torch.tensor([0.0, 0, 0, 500] * 3)
            self.cam_mins, self.cam_maxs = cams.min(0).values, cams.max(0).values

            def fmt(v):          # utils/utils.py:5-16 num_to_string: 'p' for '.', 'n' is not used by the reference ('-' kept)
                return f"{float(v):.4f}".replace(".", "p")
            self.placeholder_view_tokens = [f"<view_dtu12d_cam{k}_" + "_".join(fmt(v) for v in cams[k]) + ">" for k in range(n_views)]
        else:
            self.placeholder_view_tokens = [f"<view_{10 * (k % 3)}_{int(360 * k / n_views)}_1p2>" for k in range(n_views)]
        self.placeholder_tokens = self.placeholder_view_tokens + self.placeholder_object_tokens
        self.current_object_idx = 0
        self._rng = torch.Generator().manual_seed(seed + 1)
        self._images = torch.rand(8, 3, size[0] // 8, size[1] // 8, generator=g) * 2 - 1      # upsampled on access

    def __len__(self) -> int:
        return self._length

    def reset_sampled_object(self, idx=None) -> int:
        assert self.learnable_mode == 3
        self.current_object_idx = int(torch.randint(0, len(self.placeholder_object_tokens), (1,), generator=self._rng)) \
            if idx is None else int(idx)
        return self.current_object_idx

    def __getitem__(self, i: int):
        ex = {"image_idx": i % 8}
        img = self._images[i % 8]
        ex["pixel_values"] = torch.nn.functional.interpolate(img[None], size=self.size, mode="bilinear", align_corners=False)[0]
        obj = None
        if self.learnable_mode == 3:
            obj = self.placeholder_object_tokens[self.current_object_idx]
        elif self.learnable_mode != 1:
            obj = self.placeholder_object_tokens[0]
        view = self.placeholder_view_tokens[i % len(self.placeholder_view_tokens)] if self.placeholder_view_tokens else None
        if self.learnable_mode == 0:
            text = f"A photo of a {obj}"
        else:
            text = f"{view} . A photo of a {obj if obj is not None else self.fixed_object_token}"
        ex["text"] = text
        ex["input_ids"] = self.tokenizer(text, padding="max_length", truncation=True, max_length=self.tokenizer.model_max_length,
                                         return_tensors="pt").input_ids[0]
        ex["input_ids_placeholder_object"] = torch.tensor(self.tokenizer.convert_tokens_to_ids(obj) if obj is not None else -1)
        ex["input_ids_placeholder_view"] = torch.tensor(self.tokenizer.convert_tokens_to_ids(view) if view is not None else -1)
        return ex


def build_sd_models(device="cuda", seed: int = 0, unet_cfg=None, text_cfg: ClipEncoderConfig = SD21_TEXT, with_vae: bool = True,
                    prediction_type: str = "v_prediction"):
    """(tokenizer, noise_scheduler, text_encoder, vae, unet) with seeded weights at the SD-2.1 shapes - what
    `Coach._init_sd_models` (coach.py:600-640) returns when no checkpoint directory exists."""
    from ..models.neti_clip_text_encoder import NeTICLIPTextModel
    from ..schedulers import DDPMScheduler
    from ..sd21 import SD21, init_state_dict as unet_init
    from ..unet import UNet2DConditionModel
    unet_cfg = unet_cfg or SD21
    g = torch.Generator().manual_seed(seed)
    C = text_cfg.hidden_size
    tok = torch.randn(49408, C, generator=g) * 0.02
    pos = torch.randn(77, C, generator=g) * 0.01
    enc = CLIPEncoder(init_state_dict(text_cfg, seed), text_cfg, device)
    text_encoder = NeTICLIPTextModel.from_parts(tok, pos, (torch.ones(C), torch.zeros(C)), enc)
    unet = UNet2DConditionModel(unet_init(unet_cfg, seed), unet_cfg, device)
    vae = None
    if with_vae:
        from ..models.vae import SD21_VAE, AutoencoderKL, init_state_dict as vae_init
        vae = AutoencoderKL(vae_init(SD21_VAE, seed), SD21_VAE, device)
    return SyntheticTokenizer(), DDPMScheduler(prediction_type), text_encoder, vae, unet
