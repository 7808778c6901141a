// grx_envg.h — internal interface between the C ABI of the env (grx_env.cu) and the generic-topology env kernels (grx_phys_generic.cu).
// grx_env_create routes every model that is not the registered lower-limb tree (or every model, with GRX_ENV_GENERIC=1) here.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "grx_b200.h"
#include "grx_task.cuh"

namespace grx {
int envg_create(const grx_model_desc *md, const grx_task_cfg *cfg, int device, void **out);   // GRX_* status; *out = opaque handle
void envg_destroy(void *g);
int envg_set_self_collision(void *g, const int32_t *pairs, int32_t npairs, int32_t max_self_contacts);
int envg_launch_step(void *g, const EnvArgs &A, const grx_task_cfg &cfg, const LayR &L, bool phys, cudaStream_t st);
int envg_launch_reset(void *g, const EnvArgs &A, const grx_task_cfg &cfg, const LayR &L, const int *ids, int n, int curriculum_active, cudaStream_t st);
}  // namespace grx
