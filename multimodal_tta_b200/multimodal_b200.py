"""``unet_multimodal_deepfusion_b200`` / ``unet_multimodal_midfusion_b200`` -- drop-in for the reference's
multimodal model ``MultimodalUNetDeepFusion`` (/root/reference/src/models/unet_multimodal_midfusion.py:139-267,
configs/model/unet_multimodal_midfusion.yaml) on the B200 kernels.

Reference structure and how each piece runs here:

* one ``SpecificEncoder`` per modality (:16-78: ResidualUnits ``channels`` x ``strides + [1]`` on ONE input
  channel).  All encoders read the SAME packed 4-channel input chunk: the stem convs (1 -> C0) are launched as
  4 -> C0 convs whose weights are zero except for their own modality's channel -- ``torch.split`` is free.
* bottleneck fusion (:81-97, :216-218): ``pseudo_shared = mean_m f_m`` (tta_mean_planes), then ONE fusion layer
  (shared conv weights and norm parameters) applied to every modality: the modalities become batch entries of a
  single launch (``Act.batched_alias``: [N][M*C] and [N*M][C] are the same bytes), the residual add
  ``f_shared + Convolution(cat[f_shared, f_specific])`` is the norm kernel's residual input, and
  ``torch.cat(fused_feats, dim=1)`` for ``bottleneck_reduce`` (1x1x1, no bias) is again the same bytes.
* skips = modality means of the encoder levels (:221-224); the last decoder stage concatenates the modality-mean
  of the INPUT (:249-252), which is folded into the weights of the two convs that read it (w / M on each of the
  four input channels) -- the input needs no gradient.
* ``DecoderStage`` (:100-136): MONAI ``UpSample(mode="nontrainable")`` = 1x1x1 ``preconv`` + trilinear
  ``nn.Upsample(align_corners=True)`` (tta_upsample_fwd / _bwd), concat = channel slices, ``ResidualUnit``.
* ``final_conv`` (1x1x1) -> logits; ``domain_classifier`` (:198-199) only feeds the optional auxiliary outputs of
  the supervised trainer and is kept as parameters (state-dict compatibility), not on the adaptation path.

``state_dict()`` keys equal the reference's.  Restrictions (raised as ValueError): INSTANCE norm (the batched
fusion layer would change BatchNorm statistics), five levels (the reference hard-wires ``skip_channels_list``),
strides of 2, channel counts that are multiples of 8, <= 8 modalities.
"""
from __future__ import annotations

from typing import Any, Dict, List

import torch
import torch.nn as nn

from .config import DictConfig, create, get_config
from .registry import register_model
from .unet_b200 import B200Model, ConvHolder, ConvolutionH, ResidualUnitH


class SpecificEncoderH(nn.Module):
    def __init__(self, in_channels, channels, strides, num_res_units, act, norm, dropout):
        super().__init__()
        self.layers = nn.ModuleList()
        cur = in_channels
        for out_ch, s in zip(channels, list(strides) + [1]):
            self.layers.append(ResidualUnitH(cur, out_ch, s, 3, num_res_units, norm, act, dropout))
            cur = out_ch


class CompositionalLayerH(nn.Module):
    def __init__(self, channels, norm, act):
        super().__init__()
        self.fusion_conv = ConvolutionH(channels * 2, channels, 1, 3, norm, act, None)


class UpSampleH(nn.Module):
    """Parameter holder of MONAI UpSample("nontrainable"): ``preconv`` (1x1x1, bias) when the channels change."""

    def __init__(self, cin, cout, scale):
        super().__init__()
        self.scale = int(scale)
        if cin != cout:
            self.preconv = ConvHolder(cin, cout, 1, 1, False)
        else:
            raise ValueError("unet_multimodal_b200: UpSample without channel change is unsupported")


class DecoderStageH(nn.Module):
    def __init__(self, cin, skip, cout, stride, num_res_units, act, norm, dropout):
        super().__init__()
        self.cout, self.skip = cout, skip
        self.upsample = UpSampleH(cin, cout, stride)
        self.conv = ResidualUnitH(cout + skip, cout, 1, 3, num_res_units, norm, act, dropout)


class _BiaslessConvHolder(ConvHolder):
    """``nn.Conv3d(..., bias=False)`` (bottleneck_reduce): no bias key in the state dict."""

    def __init__(self, cin, cout, k):
        super().__init__(cin, cout, k, 1, False)
        del self.bias
        self.register_parameter("bias", None)


@register_model("unet_multimodal_deepfusion_b200")
@register_model("unet_multimodal_midfusion_b200")
class MultimodalUNetB200(B200Model):
    supports_grad_f16 = False     # its mean / upsample backward read fp32 gradient tensors

    def __init__(self, cfg: DictConfig | Dict[str, Any]):
        super().__init__()
        if not isinstance(cfg, DictConfig):
            cfg = create(dict(cfg))
        self.num_modalities = int(get_config(cfg, "num_modalities", 4))
        self.in_channels = self.num_modalities
        self.out_channels = int(get_config(cfg, "num_classes", 3))
        self.channels = [int(c) for c in get_config(cfg, "channels", [32, 64, 128, 256, 512])]
        self.strides = [int(s) for s in get_config(cfg, "strides", [2, 2, 2, 2])]
        self.num_res_units = int(get_config(cfg, "num_res_units", 2))
        self.act = get_config(cfg, "act", "RELU")
        self.norm = get_config(cfg, "norm", "INSTANCE")
        self.dropout = float(get_config(cfg, "dropout", 0.0))
        dom = get_config(cfg, "domain_classifier", create({}))
        self.domain_enabled = bool(get_config(dom, "enabled", True))
        self.domain_loss_weight = float(get_config(dom, "loss_weight", 0.1))
        if int(get_config(cfg, "spatial_dims", 3)) != 3:
            raise ValueError("unet_multimodal_b200 implements the 3-D path only")
        if str(self.norm).upper() != "INSTANCE":
            raise ValueError("unet_multimodal_b200: norm must be INSTANCE (the reference config's setting)")
        if len(self.channels) != 5 or len(self.strides) != 4:
            raise ValueError("unet_multimodal_b200: five levels / four strides (the reference hard-wires its skips)")
        if any(s != 2 for s in self.strides) or any(c % 8 for c in self.channels):
            raise ValueError("unet_multimodal_b200: strides must be 2 and channel counts multiples of 8")
        if not 1 <= self.num_modalities <= 8:
            raise ValueError("unet_multimodal_b200: 1..8 modalities")
        if self.num_res_units < 1:
            raise ValueError("unet_multimodal_b200: num_res_units >= 1")
        self._init_backend(cfg)
        ch, st, nru = self.channels, self.strides, self.num_res_units
        self.specific_encoders = nn.ModuleList([SpecificEncoderH(1, ch, st, nru, self.act, self.norm, self.dropout)
                                                for _ in range(self.num_modalities)])
        self.fusion_layer = CompositionalLayerH(ch[-1], self.norm, self.act)
        self.bottleneck_reduce = _BiaslessConvHolder(ch[-1] * self.num_modalities, ch[-1], 1)
        self.decoder_stages = nn.ModuleList()
        skip_channels = [ch[2], ch[1], ch[0], 1]
        for i in range(len(ch) - 1):
            idx = len(ch) - 1 - i
            self.decoder_stages.append(DecoderStageH(ch[idx], skip_channels[i], ch[idx - 1], st[idx - 1], nru,
                                                     self.act, self.norm, self.dropout))
        self.final_conv = ConvHolder(ch[0], self.out_channels, 1, 1, False)
        if self.domain_enabled:
            self.domain_classifier = nn.Linear(ch[-1], self.num_modalities)

    def get_domain_loss_weight(self) -> float:
        return self.domain_loss_weight if self.domain_enabled else 0.0

    # ------------------------------------------------------------------ launch options
    def configure_layers(self, engine) -> None:
        M = self.num_modalities
        Mp = (M + 7) // 8 * 8 if M > 8 else M     # modalities share ONE input chunk (M <= 8)

        def stem_map(m):
            def f(w):                              # [co][1][k,k,k] -> [co][M][k,k,k], only channel m non-zero
                out = torch.zeros((w.shape[0], Mp, *w.shape[2:]), dtype=w.dtype)
                out[:, m] = w[:, 0]
                return out
            return f

        for m, enc in enumerate(self.specific_encoders):
            first = enc.layers[0]
            u0 = first.conv.unit0.conv
            for key in (id(u0), id(first.residual)):
                cl = engine.conv_layers.get(key)
                if cl is not None:
                    cl.w_map, cl.cin = stem_map(m), Mp
            fl = engine.fused_layers.get(id(u0))
            if fl is not None:
                fl.w_map, fl.cin = stem_map(m), Mp
        # last decoder stage: input = cat[upsampled (C0), mean_m x_m (1)] -> cat[upsampled, x_0 .. x_{M-1}] with
        # the skip channel's weights divided by M on every modality channel (exact for M a power of two)
        last = self.decoder_stages[-1]
        c0 = last.cout

        def skip_map(w):                           # [co][C0 + 1][k..] -> [co][C0 + M][k..]
            out = torch.zeros((w.shape[0], c0 + Mp, *w.shape[2:]), dtype=w.dtype)
            out[:, :c0] = w[:, :c0]
            out[:, c0:c0 + M] = (w[:, c0:c0 + 1] / M).expand(-1, M, *w.shape[2:])
            return out

        for holder in (last.conv.conv.unit0.conv, last.conv.residual):
            cl = engine.conv_layers[id(holder)]
            cl.w_map, cl.cin, cl.dgrad_cin = skip_map, c0 + Mp, c0

    # ------------------------------------------------------------------ op graph
    def build_graph(self, G):
        M, ch = self.num_modalities, self.channels
        D, H, W = G.dims
        if D % 16 or H % 16 or W % 16:
            raise ValueError(f"unet_multimodal_b200: spatial size {(D, H, W)} must be divisible by 16")
        cb = ch[-1] // 8                                            # chunks of the bottleneck feature
        low = (D // 16, H // 16, W // 16)
        # [N][M * 2C]: per modality [shared | f_m]; the same bytes as [N*M][2C] feed the batched fusion conv
        catbuf = G.new_act(M * 2 * ch[-1], low, name="fusion_cat")
        skips: List[List[Any]] = []
        feats = []
        for m, enc in enumerate(self.specific_encoders):
            cur, lv = G.x.view(), []
            n_layers = len(enc.layers)
            for i, layer in enumerate(enc.layers):
                out = catbuf.view(m * 2 * cb + cb, cb) if i == n_layers - 1 else None
                cur = G.residual_unit(layer, cur, out)
                if i < n_layers - 1:
                    lv.append(cur)
            skips.append(lv)
            feats.append(cur)
        cat_b = catbuf.batched_alias(M, name="fusion_cat/batched")    # [N*M][2C]
        G.plan.keep.append(cat_b)
        G.mean(feats, cat_b.view(0, cb), rep=M)                       # pseudo-shared feature, one copy per modality
        fcv = self.fusion_layer.fusion_conv
        y = G.conv(G.conv_layer(fcv.conv), cat_b.view())
        fusedbuf = G.new_act(M * ch[-1], low, name="fused")           # torch.cat(fused_feats, dim=1)
        fused_b = fusedbuf.batched_alias(M, name="fused/batched")
        G.plan.keep.append(fused_b)
        G.normact(G.norm_layer(fcv.adn.N), y, True, cat_b.view(0, cb), fused_b.view())   # f_shared + residual
        xdec = G.cast(G.conv(G.conv_layer(self.bottleneck_reduce), fusedbuf.view()))
        dims = low
        for i, stage in enumerate(self.decoder_stages):
            s = stage.upsample.scale
            odims = (dims[0] * s, dims[1] * s, dims[2] * s)
            p = G.conv(G.conv_layer(stage.upsample.preconv), xdec)
            last = i == len(self.decoder_stages) - 1
            skip_c = self.in_channels if last else stage.skip
            cat = G.new_act(stage.cout + skip_c, odims, name=f"dec{i}_cat")
            G.upsample(p, cat.view(0, stage.cout // 8))
            sv = cat.view(stage.cout // 8, (skip_c + 7) // 8)
            if last:
                G.plan.x2 = sv                                        # second, plain-layout copy of the packed input
            else:
                G.mean([skips[m][2 - i] for m in range(M)], sv)
            xdec = G.residual_unit(stage.conv, cat.view(), None)
            dims = odims
        return G.conv(G.conv_layer(self.final_conv), xdec)
