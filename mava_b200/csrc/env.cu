// C-ABI entry points for environments, PRNG helpers and library info (include/mava_b200.h).
#include <cstring>
#include <new>

#include "env.cuh"
#include "prng.cuh"

using namespace mava;

extern "C" {

int mava_abi_version(void) { return MAVA_B200_ABI_VERSION; }

const char* mava_error_string(int code) {
  if (code == 0) return "ok";
  if (code == MAVA_E_BADARG) return "mava_b200: invalid argument";
  if (code == MAVA_E_UNSUPPORTED) return "mava_b200: configuration not supported by the kernels";
  if (code == MAVA_E_NULL) return "mava_b200: null pointer";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "mava_b200: unknown error";
}

int mava_device_info(int* out3_host) {
  MAVA_CHECK_PTR(out3_host);
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) return (int)e;
  out3_host[0] = p.multiProcessorCount;
  out3_host[1] = p.major;
  out3_host[2] = p.minor;
  return 0;
}

int mava_env_create(int kind, const void* config_host, size_t config_size, mava_env_t* out) {
  MAVA_CHECK_PTR(config_host);
  MAVA_CHECK_PTR(out);
  mava_env_s* env = new (std::nothrow) mava_env_s();
  if (!env) return MAVA_E_BADARG;
  std::memset(env, 0, sizeof(*env));
  env->kind = kind;
  int rc = MAVA_E_UNSUPPORTED;
  if (kind == MAVA_ENV_RWARE && config_size == sizeof(mava_rware_config)) {
    rc = rware_create(static_cast<const mava_rware_config*>(config_host), env);
  } else if (kind == MAVA_ENV_LBF && config_size == sizeof(mava_lbf_config)) {
    rc = lbf_create(static_cast<const mava_lbf_config*>(config_host), env);
  } else {
    rc = MAVA_E_BADARG;
  }
  if (rc != 0) {
    delete env;
    return rc;
  }
  *out = env;
  return 0;
}

int mava_env_destroy(mava_env_t env) {
  delete env;
  return 0;
}

int mava_env_dims_of(mava_env_t env, mava_env_dims* out_host) {
  MAVA_CHECK_PTR(env);
  MAVA_CHECK_PTR(out_host);
  *out_host = env->dims;
  return 0;
}

int mava_env_reset(mava_env_t env, const uint32_t* keys, uint8_t* state, int8_t* view,
                   uint8_t* mask, int num_envs, mava_stream_t s) {
  MAVA_CHECK_PTR(env);
  MAVA_CHECK_PTR(keys);
  MAVA_CHECK_PTR(state);
  MAVA_CHECK_PTR(view);
  MAVA_CHECK_PTR(mask);
  MAVA_CHECK_ARG(num_envs > 0);
  if (env->kind == MAVA_ENV_RWARE)
    return rware_reset(env, keys, state, view, mask, num_envs, as_stream(s));
  if (env->kind == MAVA_ENV_LBF)
    return lbf_reset(env, keys, state, view, mask, num_envs, as_stream(s));
  return MAVA_E_UNSUPPORTED;
}

int mava_env_step(mava_env_t env, uint8_t* state, const int8_t* action, int8_t* view,
                  uint8_t* mask, float* reward, uint8_t* done, float* ep_return,
                  int32_t* ep_length, int num_envs, int auto_reset, mava_stream_t s) {
  MAVA_CHECK_PTR(env);
  MAVA_CHECK_PTR(state);
  MAVA_CHECK_PTR(action);
  MAVA_CHECK_PTR(view);
  MAVA_CHECK_PTR(mask);
  MAVA_CHECK_PTR(reward);
  MAVA_CHECK_PTR(done);
  MAVA_CHECK_PTR(ep_return);
  MAVA_CHECK_PTR(ep_length);
  MAVA_CHECK_ARG(num_envs > 0);
  if (env->kind == MAVA_ENV_RWARE)
    return rware_step(env, state, action, view, mask, reward, done, ep_return, ep_length,
                      num_envs, auto_reset, as_stream(s));
  if (env->kind == MAVA_ENV_LBF)
    return lbf_step(env, state, action, view, mask, reward, done, ep_return, ep_length, num_envs,
                    auto_reset, as_stream(s));
  return MAVA_E_UNSUPPORTED;
}

int mava_env_peek(mava_env_t env, const uint8_t* state, int field, int32_t* out, int num_envs,
                  mava_stream_t s) {
  MAVA_CHECK_PTR(env);
  MAVA_CHECK_PTR(state);
  MAVA_CHECK_PTR(out);
  MAVA_CHECK_ARG(num_envs > 0);
  if (env->kind == MAVA_ENV_RWARE) return rware_peek(env, state, field, out, num_envs, as_stream(s));
  if (env->kind == MAVA_ENV_LBF) return lbf_peek(env, state, field, out, num_envs, as_stream(s));
  return MAVA_E_UNSUPPORTED;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// PRNG kernels
// ------------------------------------------------------------------------------------------
namespace {

__global__ void split_chain_kernel(uint32_t* key_io, uint32_t* subkeys, int n) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  Key key{key_io[0], key_io[1]};
  for (int i = 0; i < n; ++i) {
    Key sub;
    split2(key, key, sub);
    subkeys[2 * i] = sub.k0;
    subkeys[2 * i + 1] = sub.k1;
  }
  key_io[0] = key.k0;
  key_io[1] = key.k1;
}

__global__ void split_kernel(const uint32_t* key, uint32_t* out, int num) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= num) return;
  const Key k{key[0], key[1]};
  const Key o = split_n(k, (uint32_t)num, (uint32_t)j);
  out[2 * j] = o.k0;
  out[2 * j + 1] = o.k1;
}

__global__ void random_bits_kernel(const uint32_t* key, uint32_t* out, int64_t n) {
  const Key k{key[0], key[1]};
  const int64_t half = (n + 1) >> 1;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < half;
       p += (int64_t)gridDim.x * blockDim.x) {
    uint32_t lo, hi;
    random_bits_pair(k, (uint32_t)p, (uint32_t)n, lo, hi);
    out[p] = lo;
    if (p + half < n) out[p + half] = hi;
  }
}

}  // namespace

extern "C" {

int mava_prng_split_chain(uint32_t* key_io, uint32_t* subkeys, int n, mava_stream_t s) {
  MAVA_CHECK_PTR(key_io);
  MAVA_CHECK_PTR(subkeys);
  MAVA_CHECK_ARG(n > 0);
  split_chain_kernel<<<1, 32, 0, as_stream(s)>>>(key_io, subkeys, n);
  return launch_status();
}

int mava_prng_split(const uint32_t* key, uint32_t* out, int num, mava_stream_t s) {
  MAVA_CHECK_PTR(key);
  MAVA_CHECK_PTR(out);
  MAVA_CHECK_ARG(num > 0);
  split_kernel<<<ceil_div(num, 256), 256, 0, as_stream(s)>>>(key, out, num);
  return launch_status();
}

int mava_prng_random_bits(const uint32_t* key, uint32_t* out, int64_t n, mava_stream_t s) {
  MAVA_CHECK_PTR(key);
  MAVA_CHECK_PTR(out);
  MAVA_CHECK_ARG(n > 0 && n < (int64_t)0xffffffffLL);
  const int64_t half = (n + 1) >> 1;
  int blocks = (int)((half + 255) / 256);
  const int cap = sm_count() * 16;
  if (blocks > cap) blocks = cap;
  random_bits_kernel<<<blocks, 256, 0, as_stream(s)>>>(key, out, n);
  return launch_status();
}

}  // extern "C"
