"""Drop-in for TemporalRefiner (P/dvis_Plus/refiner.py:6-226): same constructor keywords, parameter names and
output dictionary.  At inference the windowed mask prediction (py:169-194) keeps mask features on the device and runs
the tcgen05 mask GEMM; the reference's host round trips (`.to(device)` py:188, `.cpu()` py:191) exist only to save
GPU memory on 16-40 GB parts and are not reproduced (180 GB HBM per B200)."""
import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from .blocks import MLP, CrossAttentionLayer, FFNLayer, SelfAttentionLayer, _fast_path, add_norm, linear
from .precision import gemm_dtype


class TemporalRefiner(nn.Module):
    def __init__(self, hidden_channel=256, feedforward_channel=2048, num_head=8, decoder_layer_num=6, mask_dim=256,
                 class_num=25, windows=5):
        super().__init__()
        self.windows = windows
        self.num_heads = num_head
        self.num_layers = decoder_layer_num
        self.transformer_obj_self_attention_layers = nn.ModuleList()
        self.transformer_time_self_attention_layers = nn.ModuleList()
        self.transformer_cross_attention_layers = nn.ModuleList()
        self.transformer_ffn_layers = nn.ModuleList()
        self.conv_short_aggregate_layers = nn.ModuleList()
        self.conv_norms = nn.ModuleList()
        for _ in range(self.num_layers):
            self.transformer_time_self_attention_layers.append(SelfAttentionLayer(hidden_channel, num_head, 0.0))
            self.conv_short_aggregate_layers.append(nn.Sequential(
                nn.Conv1d(hidden_channel, hidden_channel, kernel_size=5, stride=1, padding="same", padding_mode="replicate"),
                nn.ReLU(inplace=True),
                nn.Conv1d(hidden_channel, hidden_channel, kernel_size=3, stride=1, padding="same", padding_mode="replicate")))
            self.conv_norms.append(nn.LayerNorm(hidden_channel))
            self.transformer_obj_self_attention_layers.append(SelfAttentionLayer(hidden_channel, num_head, 0.0))
            self.transformer_cross_attention_layers.append(CrossAttentionLayer(hidden_channel, num_head, 0.0))
            self.transformer_ffn_layers.append(FFNLayer(hidden_channel, feedforward_channel, 0.0))
        self.decoder_norm = nn.LayerNorm(hidden_channel)
        self.class_embed = nn.Linear(hidden_channel, class_num + 1)
        self.mask_embed = MLP(hidden_channel, hidden_channel, mask_dim, 3)
        self.activation_proj = nn.Linear(hidden_channel, 1)

    def _short_conv(self, i, x):
        """x (bq, c, t) -> conv(k5) -> ReLU -> conv(k3), replicate padding (py:44-52,116-119)."""
        c5, c3 = self.conv_short_aggregate_layers[i][0], self.conv_short_aggregate_layers[i][2]
        dt = gemm_dtype() if _fast_path(x) else x.dtype
        x = x.to(dt)
        rep = lambda z, k: torch.cat([z[..., :1].expand(-1, -1, k), z, z[..., -1:].expand(-1, -1, k)], dim=-1)   # replicate pad
        y = F.conv1d(rep(x, 2), c5.weight.to(dt), c5.bias.to(dt))
        y = F.conv1d(rep(F.relu(y), 1), c3.weight.to(dt), c3.bias.to(dt))
        return y

    def refine(self, instance_embeds, frame_embeds):
        """The 6 refinement layers (py:104-145).  (b, c, t, q) x2 -> stacked per-layer outputs (t, l, q, b, c), fp32."""
        n_batch, n_channel, n_frames, n_instance = instance_embeds.size()
        outputs = []
        output = instance_embeds.float()
        frame_embeds = frame_embeds.float().permute(3, 0, 2, 1).flatten(1, 2)        # (q, bt, c)
        for i in range(self.num_layers):
            output = output.permute(2, 0, 3, 1).flatten(1, 2)                        # (t, bq, c)
            output = self.transformer_time_self_attention_layers[i](output, tgt_mask=None, tgt_key_padding_mask=None, query_pos=None)
            output = output.permute(1, 2, 0)                                         # (bq, c, t)
            y = self._short_conv(i, output)
            output = add_norm(self.conv_norms[i], y.transpose(1, 2).contiguous(), output.transpose(1, 2).contiguous()).transpose(1, 2)
            output = output.reshape(n_batch, n_instance, n_channel, n_frames).permute(1, 0, 3, 2).flatten(1, 2)   # (q, bt, c)
            output = self.transformer_obj_self_attention_layers[i](output, tgt_mask=None, tgt_key_padding_mask=None, query_pos=None)
            output = self.transformer_cross_attention_layers[i](output, frame_embeds, memory_mask=None,
                                                                memory_key_padding_mask=None, pos=None, query_pos=None)
            output = self.transformer_ffn_layers[i](output)
            output = output.reshape(n_instance, n_batch, n_frames, n_channel).permute(1, 3, 2, 0)                 # (b, c, t, q)
            outputs.append(output)
        return torch.stack(outputs, dim=0).permute(3, 0, 4, 1, 2)                    # (l, b, c, t, q) -> (t, l, q, b, c)

    def predict_masks(self, outputs, mask_features):
        """Masks of the last layer for a slice of frames: outputs (t', l, q, b, c) of `refine`, mask_features
        (b, t', c, h, w) of the SAME frames -> (b, q, t', h, w).  Lets every rank of a frame-sharded clip finish only
        its own frames against its local mask features."""
        dec = self.decoder_norm(outputs[:, -1:]).permute(1, 3, 0, 2, 4)              # (1, b, t', q, c)
        return self._masks(self.mask_embed(dec).float(), mask_features)[-1]

    def forward(self, instance_embeds, frame_embeds, mask_features, with_masks=True):
        """instance_embeds, frame_embeds (b, c, t, q); mask_features (b, t, c, h, w)."""
        outputs = self.refine(instance_embeds, frame_embeds)
        outputs_class, outputs_masks = self.prediction(outputs, mask_features, with_masks=with_masks)
        outputs = self.decoder_norm(outputs)
        return {
            "pred_logits": outputs_class[-1].transpose(1, 2),                        # (b, t, q, c)
            "pred_masks": None if outputs_masks is None else outputs_masks[-1],       # (b, q, t, h, w)
            "aux_outputs": self._set_aux_loss(outputs_class, outputs_masks),
            "pred_embds": outputs[:, -1].permute(2, 3, 0, 1),                        # (b, c, t, q)
        }

    @torch.jit.unused
    def _set_aux_loss(self, outputs_class, outputs_seg_masks):
        if outputs_seg_masks is None:
            return [{"pred_logits": a.transpose(1, 2)} for a in outputs_class[:-1]]
        return [{"pred_logits": a.transpose(1, 2), "pred_masks": b} for a, b in zip(outputs_class[:-1], outputs_seg_masks[:-1])]

    mask_dtype = torch.float32   # dtype of the inference-path mask logits (torch.bfloat16 halves the write traffic)

    def _masks(self, mask_embed, mask_features):
        """einsum "lbtqc,btchw->lbqthw" (py:185-189,223)."""
        if _fast_path(mask_features):
            l, b, t, q, c = mask_embed.shape
            feats = mask_features.flatten(0, 1)
            out = [ops.mask_logits(mask_embed[li].flatten(0, 1), feats, self.mask_dtype)
                   .reshape(b, t, q, *mask_features.shape[-2:]).permute(0, 2, 1, 3, 4) for li in range(l)]
            return torch.stack(out, 0)
        return torch.einsum("lbtqc,btchw->lbqthw", mask_embed.float(), mask_features.float())

    def windows_prediction(self, outputs, mask_features, windows=5, with_masks=True):
        """py:169-194, on-device: the window loop is kept (API) but every window's GEMM reads resident features."""
        iters = (outputs.size(0) + windows - 1) // windows
        outputs_classes, outputs_masks = [], []
        for i in range(iters):
            s, e = i * windows, (i + 1) * windows
            decoder_output = self.decoder_norm(outputs[s:e]).permute(1, 3, 0, 2, 4)    # (l, b, t, q, c)
            outputs_classes.append(decoder_output)
            if with_masks:
                mask_embed = self.mask_embed(decoder_output).float()
                outputs_masks.append(self._masks(mask_embed, mask_features[:, s:e].to(mask_embed.device)))
        outputs_classes = self.pred_class(torch.cat(outputs_classes, dim=2))
        return outputs_classes, (torch.cat(outputs_masks, dim=3) if with_masks else None)

    def pred_class(self, decoder_output):
        """py:196-210: softmax-over-time weighted mean of the queries, then the class head."""
        T = decoder_output.size(2)
        activation = linear(self.activation_proj, decoder_output).float().softmax(dim=2)
        class_output = (decoder_output * activation).sum(dim=2, keepdim=True).repeat(1, 1, T, 1, 1)
        return linear(self.class_embed, class_output).float().transpose(2, 3)

    def prediction(self, outputs, mask_features, with_masks=True):
        if self.training:
            decoder_output = self.decoder_norm(outputs).permute(1, 3, 0, 2, 4)
            outputs_class = self.pred_class(decoder_output)
            mask_embed = self.mask_embed(decoder_output)
            return outputs_class, torch.einsum("lbtqc,btchw->lbqthw", mask_embed.float(), mask_features.float())
        outputs = outputs[:, -1:]
        return self.windows_prediction(outputs, mask_features, windows=self.windows, with_masks=with_masks)
