"""probunet_b200: B200-native Probabilistic U-Net training / sampling hot path.

Public API mirrors the reference module `prob_unet` (ProbabilisticUNet, AxisAlignedConvGaussian, Fcomb);
everything on the device goes through the C ABI in include/probunet_b200.h (libprobunet_b200.so).
"""
__all__ = ['ProbabilisticUNet', 'AxisAlignedConvGaussian', 'Fcomb', 'UNet', 'AdamW']


def __getattr__(name):
    if name in ('ProbabilisticUNet', 'AxisAlignedConvGaussian', 'Fcomb'):
        from . import prob_unet
        return getattr(prob_unet, name)
    if name == 'UNet':
        from . import networks
        return networks.UNet
    if name == 'AdamW':
        from . import optim
        return optim.AdamW
    raise AttributeError(name)
