"""Host shell: a diffusers-protocol UNet2DConditionModel stand-in.

The reference builds its student/teacher nets with
``diffusers.UNet2DConditionModel.from_pretrained`` (train_sdxl_zh.py:138,151;
train_sd_zh.py:100,111).  diffusers is a third-party dependency that is neither
vendored under the reference tree nor installed on the build/GPU boxes, so this
file restates the *structure* of that UNet (block wiring, module names, tensor
shapes, the ``Attention``/processor protocol) so that the hot path -- the 70
(SDXL) / 16 (SD1.5) ``attn2`` cross-attention modules and the 7 / 9 KD hook
taps -- can be hosted at the reference's real shapes.

Everything here is plain PyTorch library code (cuDNN convs, cuBLAS linears,
SDPA self-attention): it is the *host* of the hot path, not the hot path.  The
pieces the product replaces are reached only through the two protocols the
reference relies on:

* ``unet.set_attn_processor(dict|proc)`` / ``unet.attn_processors`` with keys
  ``"<module path>.processor"`` (diffusers 0.23 API), and
* ``processor(attn, hidden_states, encoder_hidden_states=None,
  attention_mask=None, temb=None, scale=1.0)``.

Module/attribute names follow diffusers so that hooks written by the reference
(``unet.down_blocks[i].register_forward_hook`` -- train_sdxl_zh.py:78-84) work
unchanged.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------- config
@dataclass
class UNetConfig:
    sample_size: int = 128
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280)
    down_block_types: Tuple[str, ...] = ("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D")
    up_block_types: Tuple[str, ...] = ("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D")
    layers_per_block: int = 2
    transformer_layers_per_block: Tuple[int, ...] = (1, 2, 10)
    num_attention_heads: Tuple[int, ...] = (5, 10, 20)
    cross_attention_dim: int = 2048
    use_linear_projection: bool = True
    addition_embed_type: Optional[str] = "text_time"
    addition_time_embed_dim: int = 256
    projection_class_embeddings_input_dim: int = 2816
    norm_num_groups: int = 32

    @staticmethod
    def sdxl() -> "UNetConfig":
        """SDXL-base UNet (70 attn2 modules, cross_attention_dim 2048, head_dim 64)."""
        return UNetConfig()

    @staticmethod
    def sd15() -> "UNetConfig":
        """SD1.5 UNet (16 attn2 modules, cross_attention_dim 768, 8 heads)."""
        return UNetConfig(
            sample_size=64,
            block_out_channels=(320, 640, 1280, 1280),
            down_block_types=("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",),
            up_block_types=("UpBlock2D",) + ("CrossAttnUpBlock2D",) * 3,
            transformer_layers_per_block=(1, 1, 1, 1),
            num_attention_heads=(8, 8, 8, 8),
            cross_attention_dim=768,
            use_linear_projection=False,
            addition_embed_type=None,
        )

    @staticmethod
    def tiny(kind: str = "sdxl") -> "UNetConfig":
        """Same wiring, shrunk widths/depths -- for CPU tests that must finish in seconds."""
        if kind == "sdxl":
            return UNetConfig(
                sample_size=16, block_out_channels=(64, 128, 256), transformer_layers_per_block=(1, 1, 2),
                num_attention_heads=(1, 2, 4), cross_attention_dim=128, addition_time_embed_dim=32,
                projection_class_embeddings_input_dim=6 * 32 + 80, norm_num_groups=16,
            )
        return UNetConfig(
            sample_size=16, block_out_channels=(64, 128, 128, 128),
            down_block_types=("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",),
            up_block_types=("UpBlock2D",) + ("CrossAttnUpBlock2D",) * 3,
            transformer_layers_per_block=(1, 1, 1, 1), num_attention_heads=(2, 2, 2, 2),
            cross_attention_dim=96, use_linear_projection=False, addition_embed_type=None, norm_num_groups=16,
            layers_per_block=1,
        )


# --------------------------------------------------------------------------- attention
class StockAttnProcessor:
    """Restatement of diffusers' ``AttnProcessor2_0`` for the cases these UNets use.

    q/k/v projections, split heads, ``F.scaled_dot_product_attention`` (no mask,
    no dropout, non-causal), merge heads, ``to_out``.  This is what the
    reference runs for *every* attention module, attn2 included.
    It remains the processor for attn1 (self-attention, out of scope).
    """

    def __call__(self, attn: "Attention", hidden_states, encoder_hidden_states=None,
                 attention_mask=None, temb=None, scale: float = 1.0):
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        b, n, _ = hidden_states.shape
        q = attn.to_q(hidden_states)
        k = attn.to_k(ctx)
        v = attn.to_v(ctx)
        h = attn.heads
        d = q.shape[-1] // h
        q = q.view(b, -1, h, d).transpose(1, 2)
        k = k.view(b, -1, h, d).transpose(1, 2)
        v = v.view(b, -1, h, d).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(b, -1, h * d).to(q.dtype)
        o = attn.to_out[0](o)
        o = attn.to_out[1](o)
        return o


class Attention(nn.Module):
    """diffusers ``Attention`` attribute surface."""

    def __init__(self, query_dim: int, cross_attention_dim: Optional[int], heads: int, dim_head: int):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.inner_dim = inner
        self.is_cross_attention = cross_attention_dim is not None
        self.cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.norm_cross = None
        self.group_norm = None
        self.spatial_norm = None
        self.residual_connection = False
        self.rescale_output_factor = 1.0
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(self.cross_attention_dim, inner, bias=False)
        self.to_v = nn.Linear(self.cross_attention_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=True), nn.Dropout(0.0)])
        self.processor = StockAttnProcessor()

    def set_processor(self, processor) -> None:
        self.processor = processor

    def get_processor(self):
        return self.processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **cross_attention_kwargs)


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim: int, mult: int = 4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    """norm1 -> attn1 (self) -> +res ; norm2 -> attn2 (cross) -> +res ; norm3 -> GEGLU FF -> +res."""

    def __init__(self, dim: int, heads: int, dim_head: int, cross_attention_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, cross_attention_dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, hidden_states, encoder_hidden_states=None):
        hidden_states = self.attn1(self.norm1(hidden_states)) + hidden_states
        hidden_states = self.attn2(self.norm2(hidden_states), encoder_hidden_states=encoder_hidden_states) + hidden_states
        hidden_states = self.ff(self.norm3(hidden_states)) + hidden_states
        return hidden_states


class Transformer2DModel(nn.Module):
    def __init__(self, heads: int, dim_head: int, in_channels: int, num_layers: int, cross_attention_dim: int,
                 use_linear_projection: bool, norm_num_groups: int):
        super().__init__()
        inner = heads * dim_head
        self.use_linear_projection = use_linear_projection
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6, affine=True)
        if use_linear_projection:
            self.proj_in = nn.Linear(in_channels, inner)
            self.proj_out = nn.Linear(inner, in_channels)
        else:
            self.proj_in = nn.Conv2d(in_channels, inner, 1)
            self.proj_out = nn.Conv2d(inner, in_channels, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim) for _ in range(num_layers)])

    def forward(self, hidden_states, encoder_hidden_states=None):
        b, c, h, w = hidden_states.shape
        residual = hidden_states
        x = self.norm(hidden_states)
        if self.use_linear_projection:
            x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
            x = self.proj_in(x)
        else:
            x = self.proj_in(x)
            x = x.permute(0, 2, 3, 1).reshape(b, h * w, x.shape[1])
        for blk in self.transformer_blocks:
            x = blk(x, encoder_hidden_states=encoder_hidden_states)
        if self.use_linear_projection:
            x = self.proj_out(x)
            x = x.reshape(b, h, w, -1).permute(0, 3, 1, 2).contiguous()
        else:
            x = x.reshape(b, h, w, -1).permute(0, 3, 1, 2).contiguous()
            x = self.proj_out(x)
        return x + residual


# --------------------------------------------------------------------------- conv blocks
class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, temb_channels: int, groups: int):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=1e-5)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=1e-5)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Downsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    """DownBlock2D / CrossAttnDownBlock2D.  Returns ``(hidden_states, residual_tuple)`` --
    the KD hook stores ``output[0]`` (train_sdxl_zh.py:71-73)."""

    def __init__(self, in_ch, out_ch, temb_ch, num_layers, groups, add_downsample, attn_cfg=None):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(in_ch if i == 0 else out_ch, out_ch, temb_ch, groups) for i in range(num_layers)])
        self.has_cross_attention = attn_cfg is not None
        if attn_cfg is not None:
            self.attentions = nn.ModuleList([Transformer2DModel(in_channels=out_ch, **attn_cfg) for _ in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_ch)]) if add_downsample else None

    def forward(self, hidden_states, temb, encoder_hidden_states=None):
        outs = ()
        for i, resnet in enumerate(self.resnets):
            hidden_states = resnet(hidden_states, temb)
            if self.has_cross_attention:
                hidden_states = self.attentions[i](hidden_states, encoder_hidden_states=encoder_hidden_states)
            outs = outs + (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            outs = outs + (hidden_states,)
        return hidden_states, outs


class MidBlock(nn.Module):
    def __init__(self, ch, temb_ch, groups, attn_cfg):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb_ch, groups), ResnetBlock2D(ch, ch, temb_ch, groups)])
        self.attentions = nn.ModuleList([Transformer2DModel(in_channels=ch, **attn_cfg)])
        self.has_cross_attention = True

    def forward(self, hidden_states, temb, encoder_hidden_states=None):
        hidden_states = self.resnets[0](hidden_states, temb)
        hidden_states = self.attentions[0](hidden_states, encoder_hidden_states=encoder_hidden_states)
        return self.resnets[1](hidden_states, temb)


class UpBlock(nn.Module):
    """UpBlock2D / CrossAttnUpBlock2D."""

    def __init__(self, in_ch, out_ch, prev_out_ch, temb_ch, num_layers, groups, add_upsample, attn_cfg=None):
        super().__init__()
        resnets = []
        for i in range(num_layers):
            skip = in_ch if i == num_layers - 1 else out_ch
            rin = prev_out_ch if i == 0 else out_ch
            resnets.append(ResnetBlock2D(rin + skip, out_ch, temb_ch, groups))
        self.resnets = nn.ModuleList(resnets)
        self.has_cross_attention = attn_cfg is not None
        if attn_cfg is not None:
            self.attentions = nn.ModuleList([Transformer2DModel(in_channels=out_ch, **attn_cfg) for _ in range(num_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_ch)]) if add_upsample else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb, encoder_hidden_states=None):
        for i, resnet in enumerate(self.resnets):
            res = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res], dim=1)
            hidden_states = resnet(hidden_states, temb)
            if self.has_cross_attention:
                hidden_states = self.attentions[i](hidden_states, encoder_hidden_states=encoder_hidden_states)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states)
        return hidden_states


def sinusoidal_embedding(timesteps: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers ``Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0)``."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim: int, out_dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, out_dim)
        self.linear_2 = nn.Linear(out_dim, out_dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


# --------------------------------------------------------------------------- the UNet
class UNet2DConditionShell(nn.Module):
    """Call: ``unet(sample, timesteps, encoder_hidden_states, added_cond_kwargs=None, return_dict=False)[0]``
    exactly as train_sdxl_zh.py:397 / train_sd_zh.py:215 do."""

    def __init__(self, cfg: UNetConfig):
        super().__init__()
        self.config = cfg
        boc = cfg.block_out_channels
        temb_ch = boc[0] * 4
        g = cfg.norm_num_groups
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], temb_ch)
        if cfg.addition_embed_type == "text_time":
            self.add_embedding = TimestepEmbedding(cfg.projection_class_embeddings_input_dim, temb_ch)

        def attn_cfg(i):
            heads = cfg.num_attention_heads[i]
            return dict(heads=heads, dim_head=boc[i] // heads, num_layers=cfg.transformer_layers_per_block[i],
                        cross_attention_dim=cfg.cross_attention_dim, use_linear_projection=cfg.use_linear_projection,
                        norm_num_groups=g)

        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, kind in enumerate(cfg.down_block_types):
            in_ch, out_ch = out_ch, boc[i]
            last = i == len(boc) - 1
            self.down_blocks.append(DownBlock(in_ch, out_ch, temb_ch, cfg.layers_per_block, g, not last,
                                              attn_cfg(i) if kind.startswith("CrossAttn") else None))
        self.mid_block = MidBlock(boc[-1], temb_ch, g, attn_cfg(len(boc) - 1))
        self.up_blocks = nn.ModuleList()
        rev = tuple(reversed(boc))
        out_ch = rev[0]
        for i, kind in enumerate(cfg.up_block_types):
            prev, out_ch = out_ch, rev[i]
            in_ch = rev[min(i + 1, len(boc) - 1)]
            last = i == len(boc) - 1
            self.up_blocks.append(UpBlock(in_ch, out_ch, prev, temb_ch, cfg.layers_per_block + 1, g, not last,
                                          attn_cfg(len(boc) - 1 - i) if kind.startswith("CrossAttn") else None))
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=1e-5)
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    # ---- diffusers processor API -------------------------------------------------
    @property
    def attn_processors(self) -> Dict[str, object]:
        return {f"{name}.processor": m.processor for name, m in self.named_modules() if isinstance(m, Attention)}

    def set_attn_processor(self, processor: Union[object, Dict[str, object]]) -> None:
        attns = {f"{name}.processor": m for name, m in self.named_modules() if isinstance(m, Attention)}
        if isinstance(processor, dict):
            if len(processor) != len(attns):
                raise ValueError(
                    f"A dict of processors was passed, but the number of processors {len(processor)} does not match "
                    f"the number of attention layers: {len(attns)}. Please make sure to pass {len(attns)} processor classes.")
            for key, m in attns.items():
                m.set_processor(processor[key])
        else:
            for m in attns.values():
                m.set_processor(processor)

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    # ---- forward -----------------------------------------------------------------
    def forward(self, sample, timestep, encoder_hidden_states, added_cond_kwargs=None, return_dict: bool = False):
        cfg = self.config
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], dtype=torch.long, device=sample.device)
        timestep = timestep.expand(sample.shape[0])
        t_emb = sinusoidal_embedding(timestep, cfg.block_out_channels[0]).to(sample.dtype)
        emb = self.time_embedding(t_emb)
        if cfg.addition_embed_type == "text_time":
            text_embeds = added_cond_kwargs["text_embeds"]
            time_ids = added_cond_kwargs["time_ids"]
            tid = sinusoidal_embedding(time_ids.flatten(), cfg.addition_time_embed_dim).reshape(text_embeds.shape[0], -1)
            add = torch.cat([text_embeds, tid.to(text_embeds.dtype)], dim=-1).to(emb.dtype)
            emb = emb + self.add_embedding(add)
        sample = self.conv_in(sample)
        res = (sample,)
        for blk in self.down_blocks:
            sample, outs = blk(sample, emb, encoder_hidden_states=encoder_hidden_states)
            res = res + outs
        sample = self.mid_block(sample, emb, encoder_hidden_states=encoder_hidden_states)
        for blk in self.up_blocks:
            n = len(blk.resnets)
            sample = blk(sample, res[-n:], emb, encoder_hidden_states=encoder_hidden_states)
            res = res[:-n]
        sample = self.conv_out(F.silu(self.conv_norm_out(sample)))
        return (sample,)


def attn2_modules(unet: nn.Module):
    """(name, module) for every cross-attention (``*.attn2``) module, in forward-registration order."""
    return [(n, m) for n, m in unet.named_modules() if n.endswith(".attn2") and hasattr(m, "to_k")]
