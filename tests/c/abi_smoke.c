/* Plain C99 client of include/mava_b200.h: links libmava_b200.so and the CUDA runtime, nothing else.
 *
 *   gcc -std=c99 -Iinclude -I/usr/local/cuda/include tests/c/abi_smoke.c \
 *       -Lmava_b200 -lmava_b200 -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/mava_b200 -o abi_smoke
 *
 * Runs reset -> T env steps (fixed action stream) -> GAE through the header exactly as a foreign
 * host (the reference's FFI shim, a Go/Rust binding, ...) would, and prints FNV-1a checksums of
 * every output.  tests/test_abi_c_gpu.py runs the same sequence through the ctypes binding and
 * compares the checksums: the header, not the Python wrapper, is the contract. */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "mava_b200.h"

#define CK(call)                                                                   \
  do {                                                                             \
    int rc_ = (call);                                                              \
    if (rc_ != 0) {                                                                \
      fprintf(stderr, "%s failed: %s (%d)\n", #call, mava_error_string(rc_), rc_); \
      return 1;                                                                    \
    }                                                                              \
  } while (0)
#define CU(call)                                                            \
  do {                                                                      \
    cudaError_t e_ = (call);                                                \
    if (e_ != cudaSuccess) {                                                \
      fprintf(stderr, "%s failed: %s\n", #call, cudaGetErrorString(e_));    \
      return 1;                                                             \
    }                                                                       \
  } while (0)

static uint64_t fnv(uint64_t h, const void* p, size_t n) {
  const unsigned char* b = (const unsigned char*)p;
  for (size_t i = 0; i < n; ++i) {
    h ^= b[i];
    h *= 1099511628211ULL;
  }
  return h;
}

int main(int argc, char** argv) {
  const int E = argc > 1 ? atoi(argv[1]) : 64, T = argc > 2 ? atoi(argv[2]) : 24;
  if (mava_abi_version() != MAVA_B200_ABI_VERSION) return 2;
  mava_rware_config cfg = {8, 1, 3, /*agents*/ 2, 1, /*queue*/ 2, /*time_limit*/ 12};
  mava_env_t env = NULL;
  CK(mava_env_create(MAVA_ENV_RWARE, &cfg, sizeof(cfg), &env));
  mava_env_dims d;
  CK(mava_env_dims_of(env, &d));
  const int A = d.num_agents, FR = d.view_dim;
  cudaStream_t s;
  CU(cudaStreamCreate(&s));

  /* keys: a simple counter pattern (any uint32 pair is a valid threefry key) */
  uint32_t* hkeys = (uint32_t*)malloc(sizeof(uint32_t) * 2 * E);
  for (int e = 0; e < E; ++e) {
    hkeys[2 * e] = 0x9E3779B9u * (uint32_t)(e + 1);
    hkeys[2 * e + 1] = 0x85EBCA6Bu ^ (uint32_t)e;
  }
  int8_t* hact = (int8_t*)malloc((size_t)T * E * A);
  for (int i = 0; i < T * E * A; ++i) hact[i] = (int8_t)((i * 7 + i / 5) % 5);

  uint32_t* keys; uint8_t *state, *mask, *done; int8_t *view, *act;
  float *reward, *value, *last_val, *adv, *tgt, *ep_ret; int32_t* ep_len;
  CU(cudaMalloc((void**)&keys, sizeof(uint32_t) * 2 * E));
  CU(cudaMalloc((void**)&state, (size_t)E * d.state_stride));
  CU(cudaMalloc((void**)&view, (size_t)(T + 1) * E * A * FR));
  CU(cudaMalloc((void**)&mask, (size_t)(T + 1) * E * A));
  CU(cudaMalloc((void**)&act, (size_t)T * E * A));
  CU(cudaMalloc((void**)&reward, sizeof(float) * T * E * A));
  CU(cudaMalloc((void**)&value, sizeof(float) * T * E * A));
  CU(cudaMalloc((void**)&last_val, sizeof(float) * E * A));
  CU(cudaMalloc((void**)&adv, sizeof(float) * T * E * A));
  CU(cudaMalloc((void**)&tgt, sizeof(float) * T * E * A));
  CU(cudaMalloc((void**)&done, (size_t)T * E));
  CU(cudaMalloc((void**)&ep_ret, sizeof(float) * T * E));
  CU(cudaMalloc((void**)&ep_len, sizeof(int32_t) * T * E));
  CU(cudaMemset(state, 0, (size_t)E * d.state_stride));
  CU(cudaMemcpy(keys, hkeys, sizeof(uint32_t) * 2 * E, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(act, hact, (size_t)T * E * A, cudaMemcpyHostToDevice));
  /* a deterministic "critic": value = 0.25 * (t % 4), bootstrap 0.5 */
  float* hval = (float*)malloc(sizeof(float) * T * E * A);
  for (int t = 0; t < T; ++t)
    for (int i = 0; i < E * A; ++i) hval[t * E * A + i] = 0.25f * (float)(t % 4);
  CU(cudaMemcpy(value, hval, sizeof(float) * T * E * A, cudaMemcpyHostToDevice));
  for (int i = 0; i < E * A; ++i) hval[i] = 0.5f;
  CU(cudaMemcpy(last_val, hval, sizeof(float) * E * A, cudaMemcpyHostToDevice));

  CK(mava_env_reset(env, keys, state, view, mask, E, s));
  for (int t = 0; t < T; ++t)
    CK(mava_env_step(env, state, act + (size_t)t * E * A, view + (size_t)(t + 1) * E * A * FR,
                     mask + (size_t)(t + 1) * E * A, reward + (size_t)t * E * A, done + (size_t)t * E,
                     ep_ret + (size_t)t * E, ep_len + (size_t)t * E, E, /*auto_reset=*/1, s));
  CK(mava_gae(reward, value, done, last_val, NULL, 0.99f, 0.95f, T, E, A, 0, adv, tgt, s));
  CU(cudaStreamSynchronize(s));

#define DUMP(name, ptr, bytes)                                           \
  do {                                                                   \
    void* h_ = malloc(bytes);                                            \
    CU(cudaMemcpy(h_, ptr, bytes, cudaMemcpyDeviceToHost));              \
    printf("%s %016llx\n", name, (unsigned long long)fnv(1469598103934665603ULL, h_, bytes)); \
    free(h_);                                                            \
  } while (0)
  DUMP("state", state, (size_t)E * d.state_stride);
  DUMP("view", view, (size_t)(T + 1) * E * A * FR);
  DUMP("mask", mask, (size_t)(T + 1) * E * A);
  DUMP("reward", reward, sizeof(float) * T * E * A);
  DUMP("done", done, (size_t)T * E);
  DUMP("ep_return", ep_ret, sizeof(float) * T * E);
  DUMP("ep_length", ep_len, sizeof(int32_t) * T * E);
  DUMP("adv", adv, sizeof(float) * T * E * A);
  DUMP("targets", tgt, sizeof(float) * T * E * A);
  CK(mava_env_destroy(env));
  printf("ABI_SMOKE_OK envs %d steps %d agents %d view_dim %d\n", E, T, A, FR);
  return 0;
}
