"""``model_factory(model_params)`` -- same entry point and dispatch rule as the
reference's models/model_factory.py:25-76 ('hotformerloc' in model name)."""
from __future__ import annotations

from ..misc.utils import ModelParams
from .hotformerloc import HOTFormer, HOTFormerLoc, PoolingWrapper

_CHANNELS = {'L': 3, 'P': 3, 'D': 1, 'N': 3}


def get_in_channels(input_features: str) -> int:
    for f in input_features:
        assert f in _CHANNELS, "Invalid input features specified, must be in ['L','P','D','N']"
    n = sum(_CHANNELS[f] for f in input_features)
    assert n > 0, "Invalid input features specified, must be in ['L','P','D','N']"
    return n


def model_factory(model_params: ModelParams):
    if 'hotformerloc' not in model_params.model.lower():
        raise NotImplementedError('Model not implemented: {}'.format(model_params.model))
    mp = model_params
    backbone = HOTFormer(
        in_channels=get_in_channels(mp.input_features), channels=mp.channels,
        num_blocks=mp.num_blocks, num_heads=mp.num_heads,
        num_pyramid_levels=mp.num_pyramid_levels, num_octf_levels=mp.num_octf_levels,
        patch_size=mp.patch_size, dilation=mp.dilation, drop_path=mp.drop_path,
        stem_down=mp.num_input_downsamples, rt_size=mp.ct_size, rt_propagation=mp.ct_propagation,
        rt_propagation_scale=mp.ct_propagation_scale, disable_rt=mp.disable_rt,
        ADaPE_mode=mp.ADaPE_mode, grad_checkpoint=mp.grad_checkpoint,
        downsample_input_embeddings=mp.downsample_input_embeddings, disable_RPE=mp.disable_RPE,
        conv_norm=mp.conv_norm, layer_scale=mp.layer_scale, qkv_init=mp.qkv_init, xcpe=mp.xcpe)
    if mp.pooling == 'PyramidAttnPoolMixer' and mp.channels[-1] != 256:
        raise NotImplementedError('pooling=PyramidAttnPoolMixer needs 256 channels in the HOTFormer stage '
                                  '(hfl_attn_pool); no shipped cfg uses another width')
    pooling = PoolingWrapper(
        pool_method=mp.pooling, in_dim=mp.feature_size, output_dim=mp.output_dim,
        num_pyramid_levels=mp.num_pyramid_levels, channels=mp.channels[mp.num_octf_levels:],
        k_pooled_tokens=mp.k_pooled_tokens)
    return HOTFormerLoc(backbone=backbone, pooling=pooling,
                        normalize_embeddings=mp.normalize_embeddings,
                        input_features=mp.input_features)
