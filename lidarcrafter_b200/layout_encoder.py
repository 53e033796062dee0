"""Drop-in for the reference ``LayoutTransformerEncoder`` (lidargen/models/unets/layout_encoder.py:140-303).

Runs ONCE per ``sample()`` (diffusion/continuous_time_cond.py:268), 0.3 M parameters, 13 tokens: kept as plain
PyTorch on the device (SURVEY.md section 8 row a14 -- amortised over the 50 denoiser steps); same constructor
kwargs, same ``state_dict`` keys, same output dict.  Unlike the reference it does not call ``.cuda()`` in the
constructor (layout_encoder.py:217), so it can be built on a CPU-only host.
"""
from __future__ import annotations

import math

import torch
import torch as th
import torch.nn as nn


class LayerNorm(nn.LayerNorm):
    def forward(self, x: th.Tensor):
        return super().forward(x.float()).to(x.dtype)


class QKVMultiheadAttention(nn.Module):
    def __init__(self, n_heads: int, n_ctx: int):
        super().__init__()
        self.n_heads, self.n_ctx = n_heads, n_ctx

    def forward(self, qkv, key_padding_mask=None):
        bs, n_ctx, width = qkv.shape
        attn_ch = width // self.n_heads // 3
        scale = 1 / math.sqrt(math.sqrt(attn_ch))
        q, k, v = th.split(qkv.view(bs, n_ctx, self.n_heads, -1), attn_ch, dim=-1)
        weight = th.einsum("bthc,bshc->bhts", q * scale, k * scale)
        if key_padding_mask is not None:
            weight = weight.masked_fill(key_padding_mask.unsqueeze(1).unsqueeze(2), float("-inf"))
        weight = th.softmax(weight.float(), dim=-1).type(weight.dtype)
        return th.einsum("bhts,bshc->bthc", weight, v).reshape(bs, n_ctx, -1)


class MultiheadAttention(nn.Module):
    def __init__(self, n_ctx, width, heads):
        super().__init__()
        self.c_qkv = nn.Linear(width, width * 3)
        self.c_proj = nn.Linear(width, width)
        self.attention = QKVMultiheadAttention(heads, n_ctx)

    def forward(self, x, key_padding_mask=None):
        return self.c_proj(self.attention(self.c_qkv(x), key_padding_mask))


class MLP(nn.Module):
    def __init__(self, width):
        super().__init__()
        self.c_fc = nn.Linear(width, width * 4)
        self.c_proj = nn.Linear(width * 4, width)
        self.gelu = nn.GELU()

    def forward(self, x):
        return self.c_proj(self.gelu(self.c_fc(x)))


class ResidualAttentionBlock(nn.Module):
    def __init__(self, n_ctx: int, width: int, heads: int):
        super().__init__()
        self.attn = MultiheadAttention(n_ctx, width, heads)
        self.ln_1 = LayerNorm(width)
        self.mlp = MLP(width)
        self.ln_2 = LayerNorm(width)

    def forward(self, x, key_padding_mask=None):
        x = x + self.attn(self.ln_1(x), key_padding_mask)
        return x + self.mlp(self.ln_2(x))


class Transformer(nn.Module):
    def __init__(self, n_ctx: int, width: int, layers: int, heads: int):
        super().__init__()
        self.resblocks = nn.ModuleList([ResidualAttentionBlock(n_ctx, width, heads) for _ in range(layers)])

    def forward(self, x, key_padding_mask=None):
        for block in self.resblocks:
            x = block(x, key_padding_mask)
        return x


class LayoutTransformerEncoder(nn.Module):
    """Registry key ``"layout_encoder"``."""

    def __init__(self, feature_map_size: list, layout_length: int, hidden_dim: int, output_dim: int, num_layers: int,
                 num_heads: int, use_final_ln: bool, num_classes_for_layout_object: int,
                 mask_size_for_layout_object: int, used_condition_types=("obj_class", "obj_bbox", "obj_mask"),
                 use_positional_embedding=True, resolution_to_attention=(), use_key_padding_mask=False,
                 not_use_layout_fusion_module=False, fov_up=10, fov_down=-30, **kwargs):
        super().__init__()
        self.feature_map_size = feature_map_size
        self.not_use_layout_fusion_module = not_use_layout_fusion_module
        self.use_key_padding_mask = use_key_padding_mask
        self.used_condition_types = list(used_condition_types)
        if not not_use_layout_fusion_module:
            self.transform = Transformer(n_ctx=layout_length, width=hidden_dim, layers=num_layers, heads=num_heads)
        self.use_positional_embedding = use_positional_embedding
        if use_positional_embedding:
            self.positional_embedding = nn.Parameter(th.empty(layout_length, hidden_dim, dtype=th.float32))
            nn.init.normal_(self.positional_embedding, std=0.01)
        self.transformer_proj = nn.Linear(hidden_dim, output_dim)
        if "obj_class" in self.used_condition_types:
            self.obj_class_embedding = nn.Embedding(num_classes_for_layout_object, hidden_dim)
        if "obj_bbox" in self.used_condition_types:
            self.obj_bbox_2d_embedding = nn.Linear(4, hidden_dim)
            self.obj_bbox_embedding = nn.Linear(8, hidden_dim)
        if "obj_mask" in self.used_condition_types:
            self.obj_mask_embedding = nn.Linear(mask_size_for_layout_object * mask_size_for_layout_object, hidden_dim)
        self.final_ln = LayerNorm(hidden_dim) if use_final_ln else None
        self.dtype = torch.float32
        self.resolution_to_attention = list(resolution_to_attention)
        self._patch_tables = {}
        for r in self.resolution_to_attention:
            Hr, Wr = int(feature_map_size[0] / r), int(feature_map_size[1] / r)
            ii, ij = 1.0 / (feature_map_size[0] / r), 1.0 / (feature_map_size[1] / r)
            self._patch_tables[f"resolution{Hr}"] = torch.FloatTensor(
                [(ij * j, ii * i, ij * (j + 1), ii * (i + 1)) for i in range(Hr) for j in range(Wr)])
        self.out_channels = kwargs.get("out_channels", 10)

    def forward(self, condition_dict, obj_class=None, obj_bbox=None, obj_mask=None, is_valid_obj=None,
                image_patch_bbox=None):
        obj_bbox = condition_dict["scaled_gt_boxes"][..., :8]
        obj_bbox_2d = condition_dict["gt_boxes_2d"]
        obj_class = condition_dict["scaled_gt_boxes"][..., -1]
        is_valid_obj = condition_dict["is_valid_obj"]
        dev = obj_bbox.device
        outputs = {}
        xf_in = self.positional_embedding[None] if self.use_positional_embedding else None
        if "obj_class" in self.used_condition_types:
            e = self.obj_class_embedding(obj_class.long())
            xf_in = e if xf_in is None else xf_in + e
            outputs["obj_class_embedding"] = e.permute(0, 2, 1)
        if "obj_bbox" in self.used_condition_types:
            e3 = self.obj_bbox_embedding(obj_bbox.to(self.dtype))
            e2 = self.obj_bbox_2d_embedding(obj_bbox_2d.to(self.dtype))
            xf_in = e3 if xf_in is None else xf_in + e3 + e2
            outputs["obj_bbox_embedding"] = e2.permute(0, 2, 1)
            for key, tab in self._patch_tables.items():
                emb = self.obj_bbox_2d_embedding(tab.to(dev, self.dtype)).unsqueeze(0)
                outputs["image_patch_bbox_embedding_for_" + key] = torch.repeat_interleave(
                    emb, repeats=e3.shape[0], dim=0).permute(0, 2, 1)
        if "obj_mask" in self.used_condition_types:
            em = self.obj_mask_embedding(obj_mask.view(*obj_mask.shape[:2], -1).to(self.dtype))
            xf_in = em if xf_in is None else xf_in + em
        if "is_valid_obj" in self.used_condition_types:
            outputs["key_padding_mask"] = (1 - is_valid_obj).bool()
        kpm = outputs["key_padding_mask"] if self.use_key_padding_mask else None
        xf_out = xf_in.to(self.dtype) if self.not_use_layout_fusion_module else self.transform(xf_in.to(self.dtype), kpm)
        if self.final_ln is not None:
            xf_out = self.final_ln(xf_out)
        outputs["xf_proj"] = self.transformer_proj(xf_out[:, 0])
        outputs["xf_out"] = xf_out.permute(0, 2, 1)
        if "concat_cond" in condition_dict:
            if "autoregressive_cond" in condition_dict:
                outputs["concat_cond"] = torch.cat([condition_dict["concat_cond"],
                                                    condition_dict["autoregressive_cond"]], dim=1)
            else:
                outputs["concat_cond"] = condition_dict["concat_cond"]
        return outputs
