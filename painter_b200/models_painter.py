"""B200-native drop-in for the reference `Painter` module (Painter/models_painter.py:238-487).

Same constructor arguments, same `forward(imgs, tgts, bool_masked_pos=None, valid=None)` ->
`(loss, patchify(pred), bool_masked_pos)`, same attributes (`patch_size`, `patch_embed.num_patches`, `blocks`,
`pos_embed`, `no_weight_decay()`, `patchify/unpatchify`) and IDENTICAL state-dict keys/shapes, so
`painter_vit_large.pth` / MAE checkpoints load and `engine_train.train_one_epoch` runs unchanged.  The
`nn.Linear / nn.LayerNorm / nn.Conv2d` children are parameter holders only: their `forward` is never called —
all arithmetic goes through painter_b200.engine (hand-written sm_100a kernels).  There is no CPU path.
"""
from functools import partial

import torch
import torch.nn as nn

from . import engine, ops
from .arena import get_arena
from .engine import BlockFn, DecoderFn, EmbedFn, MergeFn, StageEnv


def trunc_normal_(tensor, mean=0.0, std=1.0):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=-2.0, b=2.0)


class Mlp(nn.Module):
    """Parameter holder with timm-0.3.2 `Mlp` key names (fc1, fc2)."""

    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden_features, in_features)


class Attention(nn.Module):
    """Parameter holder for models_painter.py:33-71 (qkv, proj, rel_pos_h, rel_pos_w)."""

    def __init__(self, dim, num_heads=8, qkv_bias=True, use_rel_pos=False, rel_pos_zero_init=True, input_size=None):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        if head_dim != 64:
            raise NotImplementedError("painter_b200: the attention kernels are specialised for head_dim = 64")
        self.scale = head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.use_rel_pos = use_rel_pos
        if not use_rel_pos:
            raise NotImplementedError("painter_b200: use_rel_pos=False is not a configuration of the reference path")
        self.rel_pos_h = nn.Parameter(torch.zeros(2 * input_size[0] - 1, head_dim))
        self.rel_pos_w = nn.Parameter(torch.zeros(2 * input_size[1] - 1, head_dim))
        if not rel_pos_zero_init:
            trunc_normal_(self.rel_pos_h, std=0.02)
            trunc_normal_(self.rel_pos_w, std=0.02)


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=True, drop_path=0.0, norm_layer=nn.LayerNorm,
                 use_rel_pos=False, rel_pos_zero_init=True, window_size=0, input_size=None):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, use_rel_pos=use_rel_pos,
                              rel_pos_zero_init=rel_pos_zero_init,
                              input_size=input_size if window_size == 0 else (window_size, window_size))
        self.drop_prob = float(drop_path)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.window_size = window_size

    def params(self):
        a, m = self.attn, self.mlp
        return (self.norm1.weight, self.norm1.bias, a.rel_pos_h, a.rel_pos_w, a.qkv.weight, a.qkv.bias,
                a.proj.weight, a.proj.bias, self.norm2.weight, self.norm2.bias, m.fc1.weight, m.fc1.bias,
                m.fc2.weight, m.fc2.bias)

    ROLES = ("n1w", "n1b", "rel_h", "rel_w", "qkv_w", "qkv_b", "proj_w", "proj_b", "n2w", "n2b", "fc1_w", "fc1_b",
             "fc2_w", "fc2_b")

    def roles(self):
        return dict(zip(self.ROLES, self.params()))


class PatchEmbed(nn.Module):
    """Parameter holder for vitdet_utils.py:160-186 (proj = Conv2d(k16, s16))."""

    def __init__(self, kernel_size=(16, 16), stride=(16, 16), padding=(0, 0), in_chans=3, embed_dim=768):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=kernel_size, stride=stride, padding=padding)


class LayerNorm2D(nn.Module):
    """Parameter holder for vitdet_utils.py:189-209."""

    def __init__(self, normalized_shape, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        self.bias = nn.Parameter(torch.zeros(normalized_shape))
        self.eps = eps


class Painter(nn.Module):
    seggpt = False
    # arithmetic of forward: "bf16" (tensor-core operands in bf16, fp32 accumulate - the training mode), "fp32"
    # (fp32-accurate split-operand mode, painter_b200/accurate.py, forward only) or "auto": fp32-accurate when the
    # module is called the way the reference's inference code calls it - outside autocast, gradients disabled
    # (seggpt_engine.py:26,47) - and bf16 otherwise (engine_train.py:65 runs under autocast)
    precision = "auto"

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=1024, depth=24, num_heads=16,
                 mlp_ratio=4., qkv_bias=True, drop_path_rate=0., norm_layer=nn.LayerNorm, act_layer=nn.GELU,
                 use_abs_pos=True, use_rel_pos=False, rel_pos_zero_init=True, window_size=0,
                 window_block_indexes=(), residual_block_indexes=(), use_act_checkpoint=False,
                 pretrain_img_size=224, pretrain_use_cls_token=True, out_feature="last_feat",
                 decoder_embed_dim=128, loss_func="smoothl1"):
        super().__init__()
        if len(tuple(residual_block_indexes)) > 0:
            raise NotImplementedError("painter_b200: residual_block_indexes (ResBottleneckBlock) is not used by any "
                                      "reference configuration and is not implemented")
        if not use_abs_pos:
            raise NotImplementedError("painter_b200: use_abs_pos=False is not implemented")
        if depth < 24:
            raise ValueError("depth must be >= 24: the taps [5, 11, 17, 23] are hard-coded (models_painter.py:416)")
        self.pretrain_use_cls_token = pretrain_use_cls_token
        self.patch_size = patch_size
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.patch_embed = PatchEmbed(kernel_size=(patch_size, patch_size), stride=(patch_size, patch_size),
                                      in_chans=in_chans, embed_dim=embed_dim)
        self.patch_embed.num_patches = (img_size[0] // patch_size) * (img_size[1] // patch_size)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, 1, embed_dim))
        self.segment_token_x = nn.Parameter(torch.zeros(1, 1, 1, embed_dim))
        self.segment_token_y = nn.Parameter(torch.zeros(1, 1, 1, embed_dim))
        self._extra_tokens(embed_dim)
        num_patches = (pretrain_img_size // patch_size) * (pretrain_img_size // patch_size)
        num_positions = (num_patches + 1) if pretrain_use_cls_token else num_patches
        self.pos_embed = nn.Parameter(torch.zeros(1, num_positions, embed_dim), requires_grad=True)

        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList()
        for i in range(depth):
            # NB: `i in window_block_indexes` is evaluated exactly like the reference (models_painter.py:307):
            # the stock factories pass a tuple of LISTS, so no block is windowed.
            self.blocks.append(Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                                     drop_path=dpr[i], norm_layer=norm_layer, use_rel_pos=use_rel_pos,
                                     rel_pos_zero_init=rel_pos_zero_init,
                                     window_size=window_size if i in window_block_indexes else 0,
                                     input_size=(img_size[0] // patch_size, img_size[1] // patch_size)))
        self._out_feature_channels = {out_feature: embed_dim}
        self._out_feature_strides = {out_feature: patch_size}
        self._out_features = [out_feature]
        trunc_normal_(self.pos_embed, std=0.02)
        self.norm = norm_layer(embed_dim)

        self.decoder_embed_dim = decoder_embed_dim
        self.decoder_embed = nn.Linear(embed_dim * 4, patch_size ** 2 * self.decoder_embed_dim, bias=True)
        self.decoder_pred = nn.Sequential(
            nn.Conv2d(self.decoder_embed_dim, self.decoder_embed_dim, kernel_size=3, padding=1),
            LayerNorm2D(self.decoder_embed_dim),
            nn.GELU(),
            nn.Conv2d(self.decoder_embed_dim, 3, kernel_size=1, bias=True),
        )
        self.loss_func = loss_func
        if loss_func not in engine.LOSS_KINDS:
            raise ValueError(f"unknown loss_func {loss_func}")
        torch.nn.init.normal_(self.mask_token, std=.02)
        torch.nn.init.normal_(self.segment_token_x, std=.02)
        torch.nn.init.normal_(self.segment_token_y, std=.02)
        self._init_extra_tokens()
        self.apply(self._init_weights)

    # ---- hooks the SegGPT subclass overrides ----
    def _extra_tokens(self, embed_dim):
        pass

    def _init_extra_tokens(self):
        pass

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'cls_token'}

    # ---- pure layout helpers (models_painter.py:355-383) ----
    def patchify(self, imgs):
        p = self.patch_size
        assert imgs.shape[2] == 2 * imgs.shape[3] and imgs.shape[2] % p == 0
        w = imgs.shape[3] // p
        h = w * 2
        x = imgs.reshape(imgs.shape[0], 3, h, p, w, p).permute(0, 2, 4, 3, 5, 1)
        return x.reshape(imgs.shape[0], h * w, p ** 2 * 3)

    def unpatchify(self, x):
        p = self.patch_size
        w = int((x.shape[1] * 0.5) ** .5)
        h = w * 2
        assert h * w == x.shape[1]
        x = x.reshape(x.shape[0], h, w, p, p, 3).permute(0, 5, 1, 3, 2, 4)
        return x.reshape(x.shape[0], 3, h * p, w * p)

    # ---- the hot path ----
    def _draw_drop_scales(self, B, device):
        """timm DropPath (models_painter.py:199,229-230) for the whole step: per-sample scales
        floor(keep + U[0,1)) / keep.  The uniform draws are made with torch.rand in the reference's order, shapes and
        dtype (attention branch then MLP branch of every block with p > 0; batch 2B up to the early merge; the
        branch output's dtype, i.e. the autocast dtype) so that a seeded run drops the same branches as the
        reference; one kernel then turns all of them into scales.  Returns {block: (scale_attn, scale_mlp)}."""
        if not self.training:
            return {}
        dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else torch.float32
        draws, keeps, slots = [], [], []
        for i, blk in enumerate(self.blocks):
            p = blk.drop_prob
            if p == 0.0:
                continue
            Bp = 2 * B if i <= 2 else B
            for br in range(2):
                draws.append(torch.rand((Bp, 1, 1, 1), dtype=dt, device=device).reshape(Bp))
                keeps.append((1.0 - p, Bp))
                slots.append((i, br, Bp))
        if not draws:
            return {}
        key = (B, str(device), tuple(k for k, _ in keeps))
        cache = self.__dict__.get("_pk_keep_cache")
        if cache is None or cache[0] != key:
            kv = torch.cat([torch.full((n,), k, dtype=torch.float32) for k, n in keeps]).to(device)
            cache = (key, kv)
            self.__dict__["_pk_keep_cache"] = cache
        scales = ops.droppath_scales(torch.cat(draws), cache[1])
        out, off = {}, 0
        for i, br, Bp in slots:
            out.setdefault(i, [None, None])[br] = scales[off:off + Bp]
            off += Bp
        return {i: tuple(v) for i, v in out.items()}

    def _drop_scales(self, i, Bp, device):
        """Per-block view of the step's DropPath scales (tests replace this to replay the reference's CPU draws)."""
        return self._step_drops.get(i, (None, None))

    def _type_emb(self, B, seg_type, device):
        return None

    def _run(self, imgs, tgts, bool_masked_pos, valid, seg_type=None, merge_between_batch=-1):
        if not imgs.is_cuda:
            raise RuntimeError("painter_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        if valid is None:
            raise ValueError("valid must be given (the reference forward_loss multiplies by it, models_painter.py:450)")
        p = self.patch_size
        B, _, H, W = imgs.shape
        h, w = H // p, W // p
        N, C = h * w, self.embed_dim
        dev = imgs.device
        imgs = imgs.detach().float().contiguous()
        tgts = tgts.detach().float().contiguous()
        valid = valid.detach().float().contiguous()
        if valid.shape != tgts.shape:
            valid = valid.expand_as(tgts).contiguous()
        mask_u8 = bool_masked_pos.reshape(-1, N).to(torch.uint8).contiguous()
        assert mask_u8.shape[0] in (1, B), "bool_masked_pos must have batch 1 or B"
        prec = self.precision
        if prec not in ("auto", "bf16", "fp32"):
            raise ValueError(f"painter_b200: precision must be 'auto', 'bf16' or 'fp32', got {prec!r}")
        if prec == "fp32" or (prec == "auto" and not torch.is_autocast_enabled("cuda") and
                              not torch.is_grad_enabled()):
            if self.training and any(b.drop_prob > 0 for b in self.blocks):
                raise NotImplementedError("painter_b200: the fp32-accurate mode is an inference mode (call .eval())")
            from . import accurate
            return accurate.forward(self, imgs, tgts, mask_u8, valid, self._type_emb(B, seg_type, dev),
                                    merge_between_batch)
        pe = self.patch_embed.proj
        # gradient arena + per-step token (engine.StageEnv): only when this forward can be followed by a backward
        arena = get_arena(self) if torch.is_grad_enabled() else None
        token = object()
        emb_prm = {"w": pe.weight, "b": pe.bias, "mt": self.mask_token, "sx": self.segment_token_x,
                   "sy": self.segment_token_y, "pe": self.pos_embed}
        z = EmbedFn.apply(imgs, tgts, mask_u8, self._type_emb(B, seg_type, dev), pe.weight, pe.bias, self.mask_token,
                          self.segment_token_x, self.segment_token_y, self.pos_embed, p, self.pretrain_use_cls_token,
                          StageEnv(arena, token, "embed", emb_prm) if arena is not None else None)
        merge_idx = 2
        Bp = 2 * B
        taps = []
        self._step_drops = self._draw_drop_scales(B, dev)
        drops = [self._drop_scales(i, 2 * B if i <= merge_idx else B, dev) for i in range(len(self.blocks))]
        for i, blk in enumerate(self.blocks):
            ws = blk.window_size
            if ws > 0 and 112 % ws != 0:
                raise NotImplementedError(f"painter_b200: window_size {ws} unsupported (must divide 112: 2,4,7,8,14,28)")
            ens_g, ens_p = 0, 0
            if merge_between_batch >= 0 and i >= merge_between_batch:
                ens_g, ens_p = (2, B) if merge_idx >= i else (1, B)
            da, dm = drops[i]
            a = blk.attn
            rel_h = engine.resize_rel_table(a.rel_pos_h, ws if ws > 0 else h)
            rel_w = engine.resize_rel_table(a.rel_pos_w, ws if ws > 0 else w)
            prm = blk.params()
            env = None
            if arena is not None:
                # hand-off: block i's dx is consumed only by block i-1's MLP branch (as bf16(DropPath scale * dx))
                # unless block i-1's output also feeds the early merge or a decoder tap
                below = None
                if i > 0 and (i - 1) != merge_idx and (i - 1) not in (5, 11, 17, 23) and ens_g == 0:
                    below = (self.blocks[i - 1].mlp.fc2.bias, drops[i - 1][1], N)
                env = StageEnv(arena, token, f"block{i}", blk.roles(), below)
            z = BlockFn.apply(z, da, dm, prm[0], prm[1], rel_h, rel_w, *prm[4:],
                              (Bp, h, w, self.num_heads, blk.norm1.eps, ens_g, ens_p, ws), env)
            if i == merge_idx:
                z = MergeFn.apply(z)
                Bp = B
            if i in (5, 11, 17, 23):
                taps.append(z)
        dp = self.decoder_pred
        dec_prm = {"norm_w": self.norm.weight, "norm_b": self.norm.bias, "dec_w": self.decoder_embed.weight,
                   "dec_b": self.decoder_embed.bias, "c3_w": dp[0].weight, "c3_b": dp[0].bias, "ln_w": dp[1].weight,
                   "ln_b": dp[1].bias, "c1_w": dp[3].weight, "c1_b": dp[3].bias}
        loss, patch = DecoderFn.apply(taps[0], taps[1], taps[2], taps[3], self.norm.weight, self.norm.bias,
                                      self.decoder_embed.weight, self.decoder_embed.bias, dp[0].weight, dp[0].bias,
                                      dp[1].weight, dp[1].bias, dp[3].weight, dp[3].bias, tgts, mask_u8, valid,
                                      (B, h, w, p, self.norm.eps, engine.LOSS_KINDS[self.loss_func], self.seggpt),
                                      StageEnv(arena, token, "decoder", dec_prm) if arena is not None else None)
        return loss.reshape(()), patch

    def forward(self, imgs, tgts, bool_masked_pos=None, valid=None):
        if bool_masked_pos is None:
            bool_masked_pos = torch.zeros((imgs.shape[0], self.patch_embed.num_patches), dtype=torch.bool,
                                          device=imgs.device)
        else:
            bool_masked_pos = bool_masked_pos.flatten(1).to(torch.bool)
        loss, pred = self._run(imgs, tgts, bool_masked_pos, valid)
        return loss, pred, bool_masked_pos


def painter_vit_large_patch16_input896x448_win_dec64_8glb_sl1(**kwargs):
    model = Painter(
        img_size=(896, 448), patch_size=16, embed_dim=1024, depth=24, num_heads=16,
        drop_path_rate=0.1, window_size=14, qkv_bias=True,
        mlp_ratio=4, norm_layer=partial(nn.LayerNorm, eps=1e-6),
        # kept verbatim from the reference factory (models_painter.py:481-482): a tuple of four lists, which
        # makes every block global-attention; the released checkpoints have the matching rel-pos shapes.
        window_block_indexes=(list(range(0, 2)) + list(range(3, 5)) + list(range(6, 8)) + list(range(9, 11)) +
                              list(range(12, 14)), list(range(15, 17)), list(range(18, 20)), list(range(21, 23))),
        residual_block_indexes=[], use_rel_pos=True, out_feature="last_feat",
        decoder_embed_dim=64,
        loss_func="smoothl1",
        **kwargs)
    return model
