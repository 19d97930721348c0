"""Deterministic U-Net baseline -- drop-in for the reference's ``baseline/deterministic_unet.py`` (:224-331).

It is the same backbone as ``networks.UNet`` with ``model_channels=64`` and self-attention switched off at every
level (the only four lines that differ in the reference), used by BASELINE.json config 5 (256x256 tiles, batch 32).
"""
from ..networks import UNet as _UNet


class UNet(_UNet):
    def __init__(self, img_resolution, in_channels, out_channels, label_dim=0, augment_dim=0, model_channels=64,
                 channel_mult=(1, 2, 3, 4), channel_mult_emb=4, num_blocks=2, attn_resolutions=(32, 16, 8),
                 dropout=0.10, label_dropout=0, use_diffuse=True):
        super().__init__(img_resolution, in_channels, out_channels, label_dim=label_dim, augment_dim=augment_dim,
                         model_channels=model_channels, channel_mult=channel_mult, channel_mult_emb=channel_mult_emb,
                         num_blocks=num_blocks, attn_resolutions=attn_resolutions, dropout=dropout,
                         label_dropout=label_dropout, use_diffuse=use_diffuse, attention=False)
