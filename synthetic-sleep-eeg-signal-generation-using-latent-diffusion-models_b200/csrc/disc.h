// Internal hooks between the autoencoder training step (engine.cu) and the discriminator (disc.cu).
#pragma once
#include <cuda_runtime.h>

struct eegldm_disc;

namespace eegldm {
// size / reset the discriminator workspace for one training step of B signals of L samples; clears its loss accumulators
int disc_prepare_step(eegldm_disc* d, int B, int L, cudaStream_t st);
// forward on the reconstruction (saved), losses[0] += MSE(act(D(recon)), 1), drecon += adv_weight * gradient
int disc_generator_term(eegldm_disc* d, const float* recon, int B, int L, float adv_weight, int no_act, float* drecon, cudaStream_t st);
// discriminator half: backward on the saved fake pass (target 0) and a fresh real pass (target 1), Adam(lr);
// losses[1] += d_fake, losses[2] += d_real
int disc_step(eegldm_disc* d, const float* x_real, int B, int L, float adv_weight, int no_act, float lr, float b1, float b2, float eps,
              cudaStream_t st);
const float* disc_losses_dev(const eegldm_disc* d);   // device [4]
}  // namespace eegldm
