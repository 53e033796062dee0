"""Layout condition encoder -- drop-in for the reference ``LayoutTransformerEncoder``
(lidargen/models/unets/layout_encoder.py:140-303): same constructor kwargs, same ``state_dict`` keys / shapes, same
output dictionary.

It runs ONCE per ``sample()`` on 13 tokens x 64 channels (0.3 M parameters), so there is nothing to win with hand-written
kernels; what matters is that it stays off the host: the module is a parameter TREE without
per-layer Python classes -- the parameters are registered under the reference's names, and ``forward`` is one
function over tensors batched across the transformer depth:

  * token embedding  = class embedding + Linear(8) of the scaled 3-D box + Linear(4) of the 2-D box (+ positional /
    mask embeddings when configured);
  * ``num_layers`` pre-LN blocks: x += W_o softmax(q k^T / sqrt(d)) v,  x += W_2 gelu(W_1 LN(x)), evaluated with
    ``F.scaled_dot_product_attention`` over a [B, heads, T, d] view of the fused QKV projection (the reference scales q and
    k by d^-1/4 each and materialises the logits; same function, one fused kernel);
  * the per-resolution patch-box embeddings depend only on the weights: they are cached per device and invalidated when
    the parameters change (``load_state_dict`` / ``_apply``).

Unlike the reference it does not call ``.cuda()`` in the constructor (layout_encoder.py:217), so it can be built on a
CPU-only host.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def _leaf(**shapes) -> nn.Module:
    """a parameter holder: nn.Module with the given tensors as Parameters (N(0, 0.02) weights, zero biases, unit LN gains)"""
    m = nn.Module()
    for name, shape in shapes.items():
        p = torch.empty(*shape)
        if name == "bias":
            nn.init.zeros_(p)
        elif len(shape) == 1:
            nn.init.ones_(p)
        else:
            nn.init.normal_(p, std=0.02)
        m.register_parameter(name, nn.Parameter(p))
    return m


def _dense(i: int, o: int) -> nn.Module:
    return _leaf(weight=(o, i), bias=(o,))


def _norm(c: int) -> nn.Module:
    return _leaf(weight=(c,), bias=(c,))


class LayoutTransformerEncoder(nn.Module):
    """Registry key ``"layout_encoder"``."""

    def __init__(self, feature_map_size: list, layout_length: int, hidden_dim: int, output_dim: int, num_layers: int,
                 num_heads: int, use_final_ln: bool, num_classes_for_layout_object: int,
                 mask_size_for_layout_object: int, used_condition_types=("obj_class", "obj_bbox", "obj_mask"),
                 use_positional_embedding=True, resolution_to_attention=(), use_key_padding_mask=False,
                 not_use_layout_fusion_module=False, fov_up=10, fov_down=-30, **kwargs):
        super().__init__()
        C = hidden_dim
        self.feature_map_size = feature_map_size
        self.layout_length, self.hidden_dim, self.num_layers, self.num_heads = layout_length, C, num_layers, num_heads
        self.not_use_layout_fusion_module = not_use_layout_fusion_module
        self.use_key_padding_mask = use_key_padding_mask
        self.used_condition_types = list(used_condition_types)
        self.use_positional_embedding = use_positional_embedding
        self.resolution_to_attention = list(resolution_to_attention)
        self.out_channels = kwargs.get("out_channels", 10)      # concat_cond channels (+ 1 for the autoregressive variant)
        self.dtype = torch.float32
        # ---- parameter tree under the reference's names ----
        if not not_use_layout_fusion_module:
            blocks = nn.ModuleList()
            for _ in range(num_layers):
                blk, attn, mlp = nn.Module(), nn.Module(), nn.Module()
                attn.add_module("c_qkv", _dense(C, 3 * C))
                attn.add_module("c_proj", _dense(C, C))
                attn.add_module("attention", nn.Module())
                mlp.add_module("c_fc", _dense(C, 4 * C))
                mlp.add_module("c_proj", _dense(4 * C, C))
                mlp.add_module("gelu", nn.Module())
                for name, sub in (("attn", attn), ("ln_1", _norm(C)), ("mlp", mlp), ("ln_2", _norm(C))):
                    blk.add_module(name, sub)
                blocks.append(blk)
            self.transform = nn.Module()
            self.transform.add_module("resblocks", blocks)
        if use_positional_embedding:
            self.positional_embedding = nn.Parameter(torch.empty(layout_length, C).normal_(std=0.01))
        self.transformer_proj = _dense(C, output_dim)
        if "obj_class" in self.used_condition_types:
            self.obj_class_embedding = _leaf(weight=(num_classes_for_layout_object, C))
        if "obj_bbox" in self.used_condition_types:
            self.obj_bbox_2d_embedding = _dense(4, C)
            self.obj_bbox_embedding = _dense(8, C)
        if "obj_mask" in self.used_condition_types:
            self.obj_mask_embedding = _dense(mask_size_for_layout_object ** 2, C)
        self.final_ln = _norm(C) if use_final_ln else None
        # normalised (x1, y1, x2, y2) of every feature-map cell at the attention resolutions (layout_encoder.py:209-217)
        self._cells = {}
        for r in self.resolution_to_attention:
            hr, wr = int(feature_map_size[0] / r), int(feature_map_size[1] / r)
            ys = torch.arange(hr, dtype=torch.float64) * (1.0 / (feature_map_size[0] / r))
            xs = torch.arange(wr, dtype=torch.float64) * (1.0 / (feature_map_size[1] / r))
            dy, dx = 1.0 / (feature_map_size[0] / r), 1.0 / (feature_map_size[1] / r)
            # the reference builds interval * index and interval * (index + 1) in Python floats, then a FloatTensor
            x1, y1 = xs[None, :].expand(hr, wr), ys[:, None].expand(hr, wr)
            x2 = (torch.arange(1, wr + 1, dtype=torch.float64) * dx)[None, :].expand(hr, wr)
            y2 = (torch.arange(1, hr + 1, dtype=torch.float64) * dy)[:, None].expand(hr, wr)
            self._cells[f"resolution{hr}"] = torch.stack([x1, y1, x2, y2], dim=-1).reshape(-1, 4).float()
        self._cell_cache = {}
        self.register_load_state_dict_post_hook(lambda mod, keys: mod._cell_cache.clear())

    def _apply(self, fn, *a, **k):
        self._cell_cache = {}
        return super()._apply(fn, *a, **k)

    # ---- forward: one function over tensors ---------------------------------------------------------------------------
    def _cell_embeddings(self, dev) -> dict:
        """Linear(4) of every cell box, [1, C, L] per attention resolution: a function of the weights only"""
        w, b = self.obj_bbox_2d_embedding.weight, self.obj_bbox_2d_embedding.bias
        key = None if w.is_inference() else (str(dev), w._version, b._version)
        if key is None or key not in self._cell_cache:
            emb = {name: F.linear(cells.to(dev), self.obj_bbox_2d_embedding.weight, self.obj_bbox_2d_embedding.bias).T[None]
                   for name, cells in self._cells.items()}
            if key is None:
                return emb
            self._cell_cache = {key: emb}
        return self._cell_cache[key]

    def _fuse(self, x: torch.Tensor, pad: torch.Tensor | None) -> torch.Tensor:
        """the layout fusion transformer on tokens [B, T, C]; pad [B, T] bool marks keys to ignore"""
        B, T, C = x.shape
        H = self.num_heads
        d = C // H
        bias = None
        if pad is not None:
            bias = torch.zeros(B, 1, 1, T, dtype=x.dtype, device=x.device).masked_fill(pad[:, None, None, :], float("-inf"))
        for blk in self.transform.resblocks:
            h = F.layer_norm(x, (C,), blk.ln_1.weight, blk.ln_1.bias)
            # fused projection, channel order [head][q | k | v][d] (layout_encoder.py:84-88)
            q, k, v = F.linear(h, blk.attn.c_qkv.weight, blk.attn.c_qkv.bias).view(B, T, H, 3, d).permute(3, 0, 2, 1, 4)
            a = F.scaled_dot_product_attention(q, k, v, attn_mask=bias, scale=1.0 / math.sqrt(d))
            x = x + F.linear(a.transpose(1, 2).reshape(B, T, C), blk.attn.c_proj.weight, blk.attn.c_proj.bias)
            h = F.layer_norm(x, (C,), blk.ln_2.weight, blk.ln_2.bias)
            x = x + F.linear(F.gelu(F.linear(h, blk.mlp.c_fc.weight, blk.mlp.c_fc.bias)), blk.mlp.c_proj.weight, blk.mlp.c_proj.bias)
        return x

    def forward(self, condition_dict, obj_class=None, obj_bbox=None, obj_mask=None, is_valid_obj=None,
                image_patch_bbox=None):
        """layout_encoder.py:237-303.  condition_dict: scaled_gt_boxes [B,T,9] (8 box numbers + class id), gt_boxes_2d [B,T,4],
        is_valid_obj [B,T] (+ concat_cond / autoregressive_cond images that are passed through)."""
        boxes = condition_dict["scaled_gt_boxes"]
        box3, cls, box2 = boxes[..., :8].to(self.dtype), boxes[..., -1].long(), condition_dict["gt_boxes_2d"].to(self.dtype)
        valid = condition_dict["is_valid_obj"]
        out, terms = {}, []
        if self.use_positional_embedding:
            terms.append(self.positional_embedding[None])
        if "obj_class" in self.used_condition_types:
            e = F.embedding(cls, self.obj_class_embedding.weight)
            out["obj_class_embedding"] = e.transpose(1, 2)
            terms.append(e)
        if "obj_bbox" in self.used_condition_types:
            e2 = F.linear(box2, self.obj_bbox_2d_embedding.weight, self.obj_bbox_2d_embedding.bias)
            out["obj_bbox_embedding"] = e2.transpose(1, 2)
            terms.append(F.linear(box3, self.obj_bbox_embedding.weight, self.obj_bbox_embedding.bias))
            if len(terms) > 1:          # the 2-D embedding joins the token only next to another term (layout_encoder.py:263-266)
                terms.append(e2)
            for name, emb in self._cell_embeddings(box3.device).items():
                out["image_patch_bbox_embedding_for_" + name] = emb.expand(box3.shape[0], -1, -1)
        if "obj_mask" in self.used_condition_types:
            terms.append(F.linear(obj_mask.flatten(2).to(self.dtype), self.obj_mask_embedding.weight, self.obj_mask_embedding.bias))
        if "is_valid_obj" in self.used_condition_types:
            out["key_padding_mask"] = (1 - valid).bool()
        tokens = terms[0]
        for t in terms[1:]:             # left-to-right like the reference's chain of additions
            tokens = tokens + t
        tokens = tokens.to(self.dtype)
        if not self.not_use_layout_fusion_module:
            tokens = self._fuse(tokens, out["key_padding_mask"] if self.use_key_padding_mask else None)
        if self.final_ln is not None:
            tokens = F.layer_norm(tokens, (self.hidden_dim,), self.final_ln.weight, self.final_ln.bias)
        out["xf_proj"] = F.linear(tokens[:, 0], self.transformer_proj.weight, self.transformer_proj.bias)
        out["xf_out"] = tokens.transpose(1, 2)
        if "concat_cond" in condition_dict:
            extra = [condition_dict["autoregressive_cond"]] if "autoregressive_cond" in condition_dict else []
            out["concat_cond"] = torch.cat([condition_dict["concat_cond"], *extra], dim=1) if extra else condition_dict["concat_cond"]
        return out
