"""The three reference configurations on the hot path, as plain dicts.

Values are those of the reference's configs/latent-diffusion/{stdiff,uncond}_cin-ldm-vq-f8.yaml
(`model.params.unet_config.params`, lines 17-36 / 14-37) and configs/autoencoder/
autoencoder_kl_32x32x4.yaml (`model.params.ddconfig`, lines 14-24).  The YAML files under the
repository's configs/ directory carry the same values with `target:` pointing at this package."""

UNET_STDIFF = dict(
    image_size=32, in_channels=4, out_channels=4, model_channels=256, attention_resolutions=[4, 2, 1],
    num_res_blocks=2, channel_mult=[1, 2, 4], num_head_channels=32, use_spatial_transformer=True,
    transformer_depth=1, context_dim=512)

UNET_UNCOND = dict(
    image_size=32, in_channels=4, out_channels=4, model_channels=256, attention_resolutions=[4, 2, 1],
    num_res_blocks=2, channel_mult=[1, 2, 4], num_head_channels=32)

AE_KL_F8_DDCONFIG = dict(
    double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
    num_res_blocks=2, attn_resolutions=[], dropout=0.0)
AE_KL_F8_EMBED_DIM = 4
# first stage of the shipped EALDM configs (configs/latent-diffusion/stdiff_cin-ldm-vq-f8.yaml:37-60)
VQ_F8_DDCONFIG = dict(
    double_z=False, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 2, 4],
    num_res_blocks=2, attn_resolutions=[32], dropout=0.0)
VQ_F8_N_EMBED, VQ_F8_EMBED_DIM = 16384, 4

DIFFUSION = dict(linear_start=0.0015, linear_end=0.0195, timesteps=1000, image_size=32, channels=4)

# algorithmic work: 2*MAC over conv / linear / attention matmuls only
UNET_STDIFF_GFLOP_PER_SAMPLE = 114.166857728
UNET_UNCOND_GFLOP_PER_SAMPLE = 79.687581696
AE_KL_DECODE_GFLOP_PER_IMAGE = 622.187
AE_KL_ENCODE_GFLOP_PER_IMAGE = 272.722
