"""Golden-vector cases for the denoiser (shared by make_golden.py and the tests)."""

CASES = {
    # name: (cfg overrides, B, T, timesteps)
    "ldm_shared_t": (dict(), 2, 768, [500]),
    "ldm_per_sample_t": (dict(), 2, 768, [980, 20]),
    "ldm_z3": (dict(in_channels=3, out_channels=3), 1, 768, [0]),
    "small_heads4": (dict(model_channels=32, channel_mult=[1, 2], attention_resolutions=[1, 2], num_heads=4, image_size=64),
                     3, 64, [7, 500, 999]),
    "small_convresample": (dict(model_channels=32, channel_mult=[1, 2, 2], attention_resolutions=[4], resblock_updown=False,
                                conv_resample=True, image_size=64), 2, 64, [123]),
    "small_poolresample": (dict(model_channels=32, channel_mult=[1, 2], attention_resolutions=[], resblock_updown=False,
                                conv_resample=False, num_res_blocks=1, image_size=48), 2, 48, [3, 4]),
    "small_headch16": (dict(model_channels=32, channel_mult=[1, 1], attention_resolutions=[2], num_head_channels=16,
                            image_size=32), 2, 32, [250.5]),
}
