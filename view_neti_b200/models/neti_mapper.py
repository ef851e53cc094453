"""NeTIMapper — the paper's mapper (arch_view_net 15) behind the reference's constructor and forward signature
(reference models/neti_mapper.py:19-47,165-169), computed by one fused CUDA kernel per direction (vn_mapper_fwd/bwd).

state_dict keys match the reference: net.{0,1,3,4}.{weight,bias}, output_layer.0.{weight,bias}; the Fourier matrix is a
buffer regenerated from seed 0 (on GPU runs of the reference it is not a Parameter either).
Supported: embedding_type "object" and "view" with (phi | theta-phi | dtu-12d) view tokens, output_bypass on/off,
norm_scale, eval-time truncation is not applied (use_nested_dropout must be False: the shipped yamls, train.yaml:21).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from .. import ops
from .._abi import VNError
from ..constants import UNET_LAYERS
from ..utils.types import MapperOutput, PESigmas
from .positional_encoding import fourier_matrix

HID = 64


def string_to_num(num: str) -> float:
    """reference utils/utils.py:19-24: inverse of num_to_string, 'p' stands for the decimal point ('1p2' -> 1.2)."""
    return float(num.replace("p", "."))


class _MapperFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, Wf, norm_scale, dim, *params):
        flat = torch.cat([p.detach().reshape(-1).float() for p in params]).contiguous()
        B = x.shape[0]
        word = torch.empty(B, dim, device=x.device)
        bypass = torch.empty(B, dim, device=x.device)
        saved = torch.empty(B, 324 + dim, device=x.device)
        ops.mapper_fwd(x.contiguous().float(), Wf, flat, norm_scale, word, bypass, saved)
        ctx.save_for_backward(flat, saved)
        ctx.norm_scale, ctx.shapes = norm_scale, [p.shape for p in params]
        return word, bypass

    @staticmethod
    def backward(ctx, d_word, d_bypass):
        flat, saved = ctx.saved_tensors
        B, dim = d_word.shape
        d_flat = torch.empty_like(flat)
        scratch = torch.empty(B, 2 * dim + HID, device=flat.device)
        ops.mapper_bwd(d_word.contiguous().float(), d_bypass.contiguous().float(), flat, saved, ctx.norm_scale, d_flat, scratch)
        grads, off = [], 0
        for shp in ctx.shapes:
            n = int(torch.Size(shp).numel())
            grads.append(d_flat[off:off + n].view(shp))
            off += n
        return (None, None, None, None, *grads)


class NeTIMapper(nn.Module):
    """ Main logic of our NeTI mapper. """

    def __init__(self, embedding_type: str, output_dim: int = 768, unet_layers: List[str] = UNET_LAYERS,
                 arch_mlp_hidden_dims: int = 128, use_nested_dropout: bool = True, nested_dropout_prob: float = 0.5,
                 norm_scale: Optional[torch.Tensor] = None, use_positional_encoding=1, num_pe_time_anchors: int = 10,
                 pe_sigmas: PESigmas = PESigmas(sigma_t=0.03, sigma_l=2.0, sigma_phi=1.0), output_bypass: bool = True,
                 placeholder_view_tokens: List[str] = None, placeholder_view_token_ids=None, arch_view_net: int = 0,
                 arch_view_mix_streams: int = 0, arch_view_disable_tl: bool = True, original_ti_init_embed=None,
                 original_ti: bool = False, bypass_unconstrained: bool = True, output_bypass_alpha: float = 0.2,
                 placeholder_object_token: str = None, cam_mins: Optional[torch.Tensor] = None,
                 cam_maxs: Optional[torch.Tensor] = None):
        super().__init__()
        if arch_view_net != 15:
            raise NotImplementedError("only arch_view_net=15 (the paper's model, neti_mapper.py:601-608) is implemented")
        if arch_view_disable_tl:
            raise NotImplementedError("For arch_view_net > 14, assume tl conditioning always")    # neti_mapper.py:483-485
        if original_ti:
            raise NotImplementedError("original_ti mappers are plain embeddings and do not use this kernel")
        if use_nested_dropout:
            raise NotImplementedError("nested dropout is off in the shipped configs (train.yaml:21) and not implemented")
        if embedding_type == "object" and arch_mlp_hidden_dims != HID:
            raise NotImplementedError("object mappers use arch_mlp_hidden_dims=64 in every shipped config (train.yaml:23); "
                                      "the fused kernel is built for that width")
        if bypass_unconstrained and not output_bypass:
            raise AssertionError("bypass_unconstrained needs output_bypass")                       # neti_mapper.py:131-132
        self.embedding_type = embedding_type
        self.arch_view_net = arch_view_net
        self.norm_scale = norm_scale
        self.output_bypass = output_bypass
        self.bypass_unconstrained = bypass_unconstrained
        self.output_bypass_alpha = output_bypass_alpha
        self.num_unet_layers = len(unet_layers)
        self.placeholder_object_token = placeholder_object_token
        self.pe_sigmas = pe_sigmas
        self.output_dim = output_dim
        sigmas = [pe_sigmas.sigma_t, pe_sigmas.sigma_l]
        if embedding_type == "view":
            self.placeholder_view_tokens = list(placeholder_view_tokens)
            self.placeholder_view_token_ids = [int(i) for i in placeholder_view_token_ids]
            self.cam_mins, self.cam_maxs = cam_mins, cam_maxs
            self._prepare_view_token_param_lookup(rescale_min_max=True)
            if self.deg_freedom == "phi":
                sigmas += [pe_sigmas.sigma_phi]
            elif self.deg_freedom == "theta-phi":
                sigmas += [pe_sigmas.sigma_theta, pe_sigmas.sigma_phi]
            else:
                sigmas += [pe_sigmas.sigma_dtu12] * 12
        elif embedding_type != "object":
            raise ValueError(embedding_type)
        self.input_dim = HID
        self.register_buffer("encoder_w", fourier_matrix(sigmas, dim=HID, seed=0), persistent=False)
        out = output_dim * 2 if output_bypass else output_dim
        self.net = nn.Sequential(nn.Linear(HID, HID), nn.LayerNorm(HID), nn.LeakyReLU(),
                                 nn.Linear(HID, HID), nn.LayerNorm(HID), nn.LeakyReLU())
        self.output_layer = nn.Sequential(nn.Linear(HID, out))
        self.name = placeholder_object_token if embedding_type == "object" else "view"

    # ---- view tokens (reference neti_mapper.py:208-280, 294-337, 440-468) -----------------------------------------
    def _prepare_view_token_param_lookup(self, rescale_min_max: bool = False):
        assert len(self.placeholder_view_tokens) == len(self.placeholder_view_token_ids)
        if "dtu12d" not in self.placeholder_view_tokens[0]:
            assert all(s[:6] == "<view_" for s in self.placeholder_view_tokens), "not view tokens"
            params = [[string_to_num(n) for n in tok[6:-1].split("_")] for tok in self.placeholder_view_tokens]
            self.view_tokenid_2_view_params = dict(zip(self.placeholder_view_token_ids, params))
            if rescale_min_max:
                allp = torch.tensor(params)
                self.theta_min, self.theta_max = allp[:, 0].min().item(), allp[:, 0].max().item()
                self.phi_min, self.phi_max = allp[:, 1].min().item(), allp[:, 1].max().item()
                self.r_min, self.r_max = allp[:, 2].min().item(), allp[:, 2].max().item()
                self.deg_freedom = "phi" if self.theta_min - self.theta_max == 0 else "theta-phi"
        else:
            self.deg_freedom = "dtu-12d"
            self.view_tokenid_2_view_params = {}
            for tok, tid in zip(self.placeholder_view_tokens, self.placeholder_view_token_ids):
                self.view_tokenid_2_view_params[tid] = torch.tensor([string_to_num(n) for n in tok[:-1].split("_")[3:]])
            if rescale_min_max and (self.cam_mins is None or self.cam_maxs is None):
                raise VNError("dtu-12d view mappers need cam_mins / cam_maxs (12-vectors over all DTU cameras; the "
                              "reference reads them from the DTU calibration files, neti_mapper.py:269-280)")

    def add_view_tokens_to_vocab(self, placeholder_view_tokens_new: List[str], placeholder_view_token_ids_new: List[int]):
        assert len(placeholder_view_tokens_new) == len(placeholder_view_token_ids_new)
        for tok, tid in zip(placeholder_view_tokens_new, placeholder_view_token_ids_new):
            if tok not in self.placeholder_view_tokens:
                self.placeholder_view_tokens.append(tok)
                self.placeholder_view_token_ids.append(int(tid))
        self._prepare_view_token_param_lookup(rescale_min_max=False)

    @staticmethod
    def scale_m1_1(x, xmin, xmax):
        if not torch.is_tensor(xmin) and xmin == xmax:
            return x
        return (x - xmin) / (xmax - xmin) * 2 - 1

    # ---- forward -----------------------------------------------------------------------------------------------------
    def _encode_inputs(self, timestep, unet_layer, input_ids_placeholder_view) -> torch.Tensor:
        """reference do_positional_encoding (:542-562): (t, l[, view]) scaled to [-1, 1], one row per sample."""
        dev = self.encoder_w.device
        t = timestep.to(dev).float() / 1000 * 2 - 1
        l = unet_layer.to(dev).float() / self.num_unet_layers * 2 - 1
        cols = [t, l]
        if self.embedding_type == "view":
            cols += self._view_columns(tuple(int(i) for i in input_ids_placeholder_view), dev)
        return torch.stack(cols, dim=1).contiguous()

    def _view_columns(self, ids: tuple, dev) -> list:
        """Scaled camera parameters of the rows' view tokens as device columns.  Cached by the id tuple: they are built from
        host numbers, and a host -> device copy from pageable memory drains the stream first (a stall per step otherwise)."""
        cache = self.__dict__.setdefault("_view_cache", {})
        key = (ids, str(dev), len(self.placeholder_view_token_ids))
        cols = cache.get(key)
        if cols is not None:
            return cols
        if len(cache) > 4096:
            cache.clear()
        vp = [self.view_tokenid_2_view_params[i] for i in ids]
        if self.deg_freedom in ("phi", "theta-phi"):
            th = self.scale_m1_1(torch.tensor([v[0] for v in vp], device=dev), self.theta_min, self.theta_max)
            ph = self.scale_m1_1(torch.tensor([v[1] for v in vp], device=dev), self.phi_min, self.phi_max)
            cols = [ph] if self.deg_freedom == "phi" else [th, ph]
        else:
            cam = torch.stack(vp).to(dev).float()
            cam = self.scale_m1_1(cam, self.cam_mins.to(dev), self.cam_maxs.to(dev))
            cols = [c.contiguous() for c in cam.unbind(1)]
        cache[key] = cols
        return cols

    def forward(self, timestep: torch.Tensor, unet_layer: torch.Tensor, input_ids_placeholder_view: torch.Tensor,
                truncation_idx: int = None) -> MapperOutput:
        if not self.encoder_w.is_cuda:
            raise VNError("NeTIMapper runs on the CUDA library only (move it with .cuda(); there is no CPU fallback)")
        x = self._encode_inputs(timestep, unet_layer, input_ids_placeholder_view)
        ns = float(self.norm_scale) if self.norm_scale is not None else 0.0
        dim = self.output_dim
        if not self.output_bypass:
            raise NotImplementedError("output_bypass=False mappers are not implemented (every shipped config uses bypass)")
        params = [self.net[0].weight, self.net[0].bias, self.net[1].weight, self.net[1].bias, self.net[3].weight,
                  self.net[3].bias, self.net[4].weight, self.net[4].bias, self.output_layer[0].weight,
                  self.output_layer[0].bias]
        word, bypass = _MapperFn.apply(x, self.encoder_w, ns, dim, *params)
        return MapperOutput(word_embedding=word, bypass_output=bypass, bypass_unconstrained=self.bypass_unconstrained,
                            output_bypass_alpha=self.output_bypass_alpha)
