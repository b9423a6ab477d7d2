"""Golden-vector cases of the latent-diffusion training step (shared by make_golden_ldm_train.py and the tests)."""

TRAIN_CASES = {
    # name: (cfg overrides, B, T, prediction_type, schedule, (beta_start, beta_end))      train_ldm.py:199-200 / sample_trials.py:136-143
    "small_eps": (dict(model_channels=32, channel_mult=[1, 2], attention_resolutions=[1, 2], num_heads=4, image_size=64),
                  3, 64, "epsilon", "linear_beta", (0.0015, 0.0195)),
    "small_v_convresample": (dict(model_channels=32, channel_mult=[1, 2, 2], attention_resolutions=[4], resblock_updown=False,
                                  conv_resample=True, image_size=64), 2, 64, "v_prediction", "scaled_linear_beta", (0.0015, 0.0205)),
    "small_poolresample_z3": (dict(model_channels=32, channel_mult=[1, 2], attention_resolutions=[2], resblock_updown=False,
                                   conv_resample=False, num_res_blocks=1, image_size=48, in_channels=3, out_channels=3),
                              2, 48, "epsilon", "linear_beta", (0.0015, 0.0195)),
    "ldm_full": (dict(), 2, 768, "epsilon", "linear_beta", (0.0015, 0.0195)),
}
FULL_GRADS = ("small_eps", "small_poolresample_z3")   # every gradient stored; the other cases: digests
