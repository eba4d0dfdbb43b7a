// Internal env handle shared by the per-environment translation units.
#pragma once
#include "common.cuh"

namespace mava {

constexpr int kMaxCells = 1024;   // H*W
constexpr int kMaxShelves = 254;  // shelf ids are stored +1 in a uint8 grid
constexpr int kMaxAgents = 8;
constexpr int kMaxQueue = 8;

// Immutable RobotWarehouse scenario constants, passed to kernels by value.
// Packed per-env record (bytes, `stride` per env, stride/16 odd so that the records of the envs of
// one warp start in different shared-memory banks):
//   off_agents   A words   x | y << 8 | direction << 16 | carrying << 24
//   off_queue    Q bytes   requested shelf ids (request_queue)
//   off_reqbits  ceil(n/32) words, bit s = shelf s is requested
//   off_step u32, off_key 2 x u32 (State.key), off_mkey 2 x u32 (RecordEpisodeMetrics key),
//   off_run_ret f32, off_run_len i32, off_ep_ret f32, off_ep_len i32
//   off_cells    HW bytes  the SHELVES grid channel: shelf id + 1 per cell, 0 = empty
// The AGENTS grid channel is never stored: with agents on distinct cells it is a function of the
// agent positions (see env_rware.cu).
struct RwareConst {
  int H, W, HW, A, Q, n, R, FR, time_limit;
  int stride;  // bytes per env record in HBM (multiple of 16)
  int off_agents, off_queue, off_reqbits;
  int off_step, off_key, off_mkey, off_run_ret, off_run_len, off_ep_ret, off_ep_len;
  int off_cells, cells_words, req_words;
  int goal[2];                            // flat cell index, scan order
  uint32_t highway[kMaxCells / 32];       // bit per cell
  uint16_t shelf_home[kMaxShelves + 2];   // flat cell index of shelf s at reset
};

struct LbfConst {
  int S, fov, A, NF, max_level, force_coop, time_limit, individual_rewards, FR;
  int stride;
  int off_ax, off_ay, off_alvl, off_fx, off_fy, off_flvl, off_featen;
  int off_step, off_key, off_mkey, off_run_ret, off_run_len, off_ep_ret, off_ep_len;
  uint32_t interior[8];  // bit per cell (grid up to 16 x 16): not on the border (food may spawn)
  uint32_t allcells[8];  // bit per cell of the grid
};

}  // namespace mava

struct mava_env_s {
  int kind;
  mava_env_dims dims;
  mava::RwareConst rw;
  mava::LbfConst lbf;
};

namespace mava {
int rware_create(const mava_rware_config* cfg, mava_env_s* env);
int rware_reset(const mava_env_s* env, const uint32_t* keys, uint8_t* state, int8_t* view,
                uint8_t* mask, int num_envs, cudaStream_t s);
int rware_step(const mava_env_s* env, uint8_t* state, const int8_t* action, int8_t* view,
               uint8_t* mask, float* reward, uint8_t* done, float* ep_return, int32_t* ep_length,
               int num_envs, int auto_reset, cudaStream_t s);
int rware_peek(const mava_env_s* env, const uint8_t* state, int field, int32_t* out, int num_envs,
               cudaStream_t s);

int lbf_create(const mava_lbf_config* cfg, mava_env_s* env);
int lbf_reset(const mava_env_s* env, const uint32_t* keys, uint8_t* state, int8_t* view,
              uint8_t* mask, int num_envs, cudaStream_t s);
int lbf_step(const mava_env_s* env, uint8_t* state, const int8_t* action, int8_t* view,
             uint8_t* mask, float* reward, uint8_t* done, float* ep_return, int32_t* ep_length,
             int num_envs, int auto_reset, cudaStream_t s);
int lbf_peek(const mava_env_s* env, const uint8_t* state, int field, int32_t* out, int num_envs,
             cudaStream_t s);
}  // namespace mava
