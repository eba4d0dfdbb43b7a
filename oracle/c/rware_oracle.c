/* Oracle, C port: Jumanji RobotWarehouse + the Mava wrapper stack, batched over envs (OpenMP).
 *
 * TEST INFRASTRUCTURE ONLY.  A plain-C restatement of oracle/rware.py (which cites the reference
 * lines it follows: mava/wrappers/jumanji.py:135-144, auto_reset_wrapper.py:60-101,
 * episode_metrics.py:59-111, make_env.py:69-83; inner env = published Jumanji algorithm, parity
 * unpinned).  Used for parity at sizes the numpy oracle cannot reach and as the CPU baseline
 * (`cpu_baseline.kind = "port"`).  tests/test_oracle_cpu.py checks it against oracle/rware.py.
 *
 * State per env (int32 words): see rw_state_words().  Plain, unpacked, nothing shared with the
 * CUDA layout.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int H, W, A, Q, R, n, FR, time_limit;
  int column_height;
} rw_spec;

static inline uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

static void threefry(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t* o0,
                     uint32_t* o1) {
  static const int rot[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
  uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  x0 += ks[0];
  x1 += ks[1];
  for (int r = 0; r < 5; ++r) {
    for (int i = 0; i < 4; ++i) {
      x0 += x1;
      x1 = rotl(x1, rot[r % 2][i]);
      x1 ^= x0;
    }
    x0 += ks[(r + 1) % 3];
    x1 += ks[(r + 2) % 3] + (uint32_t)(r + 1);
  }
  *o0 = x0;
  *o1 = x1;
}

/* jax.random.split(key) -> (a, b) */
static void split2(const uint32_t* k, uint32_t* a, uint32_t* b) {
  uint32_t y00, y01, y10, y11;
  threefry(k[0], k[1], 0, 2, &y00, &y10);
  threefry(k[0], k[1], 1, 3, &y01, &y11);
  a[0] = y00; a[1] = y01;
  b[0] = y10; b[1] = y11;
}

/* jax random_bits(key, (size,)) */
static void random_bits(const uint32_t* k, int size, uint32_t* out) {
  int half = (size + 1) / 2;
  for (int p = 0; p < half; ++p) {
    uint32_t c1 = (p + half < size) ? (uint32_t)(p + half) : 0u, lo, hi;
    threefry(k[0], k[1], (uint32_t)p, c1, &lo, &hi);
    out[p] = lo;
    if (p + half < size) out[p + half] = hi;
  }
}

/* first `num` entries of jax.random.permutation-style shuffle of vals[0..size) (one round) */
static void shuffle_first(const uint32_t* key, const int* vals, int size, int num, int* out) {
  uint32_t k2[2], sub[2];
  uint32_t* bits = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)size);
  char* used = (char*)calloc((size_t)size, 1);
  split2(key, k2, sub);
  random_bits(sub, size, bits);
  for (int r = 0; r < num; ++r) { /* selection = prefix of a stable sort */
    int best = -1;
    for (int i = 0; i < size; ++i)
      if (!used[i] && (best < 0 || bits[i] < bits[best])) best = i;
    used[best] = 1;
    out[r] = vals[best];
  }
  free(bits);
  free(used);
}

static int is_highway(const rw_spec* s, int r, int c) {
  return (c % 3 == 0) || (r % (s->column_height + 1) == 0) || (r == s->H - 1) ||
         ((r > s->H - (s->column_height + 3)) && (c == s->W / 2 - 1 || c == s->W / 2));
}

void rw_make_spec(int column_height, int shelf_rows, int shelf_columns, int num_agents,
                  int sensor_range, int queue, int time_limit, rw_spec* s) {
  s->column_height = column_height;
  s->H = (column_height + 1) * shelf_rows + 2;
  s->W = 3 * shelf_columns + 1;
  s->A = num_agents;
  s->Q = queue;
  s->R = sensor_range;
  s->time_limit = time_limit;
  int loc = (2 * sensor_range + 1) * (2 * sensor_range + 1);
  s->FR = 8 + (loc - 1) * 5 + loc * 2;
  s->n = 0;
  for (int r = 0; r < s->H; ++r)
    for (int c = 0; c < s->W; ++c)
      if (!is_highway(s, r, c)) s->n++;
}

/* state words: ax[A] ay[A] dir[A] carry[A] sx[n] sy[n] req[n] queue[Q] step key[2] mkey[2]
 *              run_len ep_len | run_ret ep_ret (floats bit-cast) */
int rw_state_words(const rw_spec* s) { return 4 * s->A + 3 * s->n + s->Q + 1 + 2 + 2 + 2 + 2; }

typedef struct {
  int *ax, *ay, *dir, *carry, *sx, *sy, *req, *queue, *step;
  uint32_t *key, *mkey;
  int *run_len, *ep_len;
  float *run_ret, *ep_ret;
} rw_view;

static rw_view view_of(const rw_spec* s, int32_t* w) {
  rw_view v;
  v.ax = w; w += s->A;
  v.ay = w; w += s->A;
  v.dir = w; w += s->A;
  v.carry = w; w += s->A;
  v.sx = w; w += s->n;
  v.sy = w; w += s->n;
  v.req = w; w += s->n;
  v.queue = w; w += s->Q;
  v.step = w; w += 1;
  v.key = (uint32_t*)w; w += 2;
  v.mkey = (uint32_t*)w; w += 2;
  v.run_len = w; w += 1;
  v.ep_len = w; w += 1;
  v.run_ret = (float*)w; w += 1;
  v.ep_ret = (float*)w;
  return v;
}

static void forward(const rw_spec* s, int x, int y, int d, int* nx, int* ny) {
  *nx = x;
  *ny = y;
  if (d == 0) *nx = x > 0 ? x - 1 : 0;
  else if (d == 1) *ny = y < s->W - 1 ? y + 1 : s->W - 1;
  else if (d == 2) *nx = x < s->H - 1 ? x + 1 : s->H - 1;
  else *ny = y > 0 ? y - 1 : 0;
}

static void build_grids(const rw_spec* s, const rw_view* v, int* gsh, int* gag) {
  memset(gsh, 0, sizeof(int) * (size_t)(s->H * s->W));
  memset(gag, 0, sizeof(int) * (size_t)(s->H * s->W));
  for (int i = 0; i < s->n; ++i) gsh[v->sx[i] * s->W + v->sy[i]] = i + 1;
  for (int i = 0; i < s->A; ++i) gag[v->ax[i] * s->W + v->ay[i]] = i + 1;
}

static void generate(const rw_spec* s, rw_view* v, const uint32_t* key_in) {
  uint32_t key[2] = {key_in[0], key_in[1]}, sub[2], d1[2], d2[2];
  int HW = s->H * s->W;
  int* cells = (int*)malloc(sizeof(int) * (size_t)HW);
  int* ids = (int*)malloc(sizeof(int) * (size_t)s->n);
  int pick[64];
  for (int i = 0; i < HW; ++i) cells[i] = i;
  split2(key, key, sub);
  shuffle_first(sub, cells, HW, s->A, pick);
  for (int i = 0; i < s->A; ++i) {
    v->ax[i] = pick[i] / s->W;
    v->ay[i] = pick[i] % s->W;
    v->carry[i] = 0;
  }
  split2(key, key, sub);
  split2(sub, d1, d2);
  {
    uint32_t bits[64];
    random_bits(d2, s->A, bits);
    for (int i = 0; i < s->A; ++i) v->dir[i] = (int)(bits[i] % 4u);
  }
  split2(key, key, sub);
  for (int i = 0; i < s->n; ++i) ids[i] = i;
  shuffle_first(sub, ids, s->n, s->Q, pick);
  int k = 0;
  for (int r = 0; r < s->H; ++r)
    for (int c = 0; c < s->W; ++c)
      if (!is_highway(s, r, c)) {
        v->sx[k] = r;
        v->sy[k] = c;
        v->req[k] = 0;
        ++k;
      }
  for (int i = 0; i < s->Q; ++i) {
    v->queue[i] = pick[i];
    v->req[pick[i]] = 1;
  }
  *v->step = 0;
  v->key[0] = key[0];
  v->key[1] = key[1];
  free(cells);
  free(ids);
}

static void observe(const rw_spec* s, const rw_view* v, const int* gsh, const int* gag,
                    int8_t* view, uint8_t* mask) {
  int loc = (2 * s->R + 1) * (2 * s->R + 1);
  for (int i = 0; i < s->A; ++i) {
    int8_t* o = view + (size_t)i * s->FR;
    int x = v->ax[i], y = v->ay[i], d = v->dir[i];
    o[0] = (int8_t)x; o[1] = (int8_t)y; o[2] = (int8_t)v->carry[i];
    for (int k = 0; k < 4; ++k) o[3 + k] = d == k;
    o[7] = (int8_t)is_highway(s, x, y);
    int ia = 8, is = 8 + (loc - 1) * 5;
    for (int dx = -s->R; dx <= s->R; ++dx)
      for (int dy = -s->R; dy <= s->R; ++dy) {
        int cx = x + dx, cy = y + dy;
        int inside = cx >= 0 && cx < s->H && cy >= 0 && cy < s->W;
        int aid = inside ? gag[cx * s->W + cy] : 0, sid = inside ? gsh[cx * s->W + cy] : 0;
        if (dx || dy) {
          o[ia] = aid != 0;
          for (int k = 0; k < 4; ++k) o[ia + 1 + k] = aid ? (v->dir[aid - 1] == k) : 0;
          ia += 5;
        }
        o[is] = sid != 0;
        o[is + 1] = sid ? (int8_t)v->req[sid - 1] : 0;
        is += 2;
      }
    int nx, ny;
    forward(s, x, y, d, &nx, &ny);
    int bad = (nx == x && ny == y) || (v->carry[i] && gsh[nx * s->W + ny] != 0);
    mask[i] = (uint8_t)(0x1D | (bad ? 0 : 2));
  }
}

void rw_reset(const rw_spec* s, const uint32_t* keys, int32_t* state, int8_t* view, uint8_t* mask,
              int num_envs) {
  int words = rw_state_words(s);
#pragma omp parallel
  {
    int* gsh = (int*)malloc(sizeof(int) * (size_t)(s->H * s->W));
    int* gag = (int*)malloc(sizeof(int) * (size_t)(s->H * s->W));
#pragma omp for schedule(static)
    for (int e = 0; e < num_envs; ++e) {
      int32_t* w = state + (size_t)e * words;
      memset(w, 0, sizeof(int32_t) * (size_t)words);
      rw_view v = view_of(s, w);
      uint32_t key[2], rkey[2];
      split2(keys + 2 * (size_t)e, key, rkey);
      generate(s, &v, rkey);
      v.mkey[0] = key[0];
      v.mkey[1] = key[1];
      build_grids(s, &v, gsh, gag);
      observe(s, &v, gsh, gag, view + (size_t)e * s->A * s->FR, mask + (size_t)e * s->A);
    }
    free(gsh);
    free(gag);
  }
}

void rw_step(const rw_spec* s, int32_t* state, const int8_t* action, int8_t* view, uint8_t* mask,
             float* reward, uint8_t* done, float* ep_return, int32_t* ep_length, int num_envs,
             int auto_reset) {
  int words = rw_state_words(s);
#pragma omp parallel
  {
    int* gsh = (int*)malloc(sizeof(int) * (size_t)(s->H * s->W));
    int* gag = (int*)malloc(sizeof(int) * (size_t)(s->H * s->W));
    int* niq = (int*)malloc(sizeof(int) * (size_t)s->n);
#pragma omp for schedule(static)
    for (int e = 0; e < num_envs; ++e) {
      rw_view v = view_of(s, state + (size_t)e * words);
      build_grids(s, &v, gsh, gag);
      int act[64];
      for (int i = 0; i < s->A; ++i) {
        int a = action[(size_t)e * s->A + i];
        if (a == 1) {
          int nx, ny;
          forward(s, v.ax[i], v.ay[i], v.dir[i], &nx, &ny);
          if ((nx == v.ax[i] && ny == v.ay[i]) || (v.carry[i] && gsh[nx * s->W + ny] != 0)) a = 0;
        }
        act[i] = a;
      }
      for (int i = 0; i < s->A; ++i) {
        int x = v.ax[i], y = v.ay[i], cell = x * s->W + y;
        if (act[i] == 2) v.dir[i] = (v.dir[i] + 3) % 4;
        else if (act[i] == 3) v.dir[i] = (v.dir[i] + 1) % 4;
        else if (act[i] == 1) {
          int nx, ny;
          forward(s, x, y, v.dir[i], &nx, &ny);
          int ncell = nx * s->W + ny;
          gag[cell] = 0;
          gag[ncell] = i + 1;
          if (v.carry[i]) {
            int sid = gsh[cell], k = sid ? sid - 1 : s->n - 1;
            v.sx[k] = nx;
            v.sy[k] = ny;
            gsh[cell] = 0;
            gsh[ncell] = sid;
          }
          v.ax[i] = nx;
          v.ay[i] = ny;
        } else if (act[i] == 4) {
          int sid = gsh[cell];
          if (!v.carry[i]) { if (sid) v.carry[i] = 1; }
          else if (!is_highway(s, x, y)) v.carry[i] = 0;
        }
      }
      int collision = 0;
      for (int i = 0; i < s->A; ++i)
        if (gag[v.ax[i] * s->W + v.ay[i]] != i + 1) collision = 1;
      float rew = 0.0f;
      for (int g = 0; g < 2; ++g) {
        int gc = (s->H - 1) * s->W + s->W / 2 - 1 + g;
        int sid = gsh[gc];
        if (sid && v.req[sid - 1] == 1) {
          uint32_t rkey[2];
          split2(v.key, v.key, rkey);
          int m = 0, pick;
          for (int k = 0; k < s->n; ++k) {
            int inq = 0;
            for (int q = 0; q < s->Q; ++q) inq |= v.queue[q] == k;
            if (!inq) niq[m++] = k;
          }
          shuffle_first(rkey, niq, m, 1, &pick);
          for (int q = 0; q < s->Q; ++q)
            if (v.queue[q] == sid - 1) { v.queue[q] = pick; break; }
          v.req[sid - 1] = 0;
          v.req[pick] = 1;
          rew += 1.0f;
        }
      }
      *v.step += 1;
      int is_done = collision || *v.step >= s->time_limit;
      float new_ret = *v.run_ret + rew;
      int new_len = *v.run_len + 1;
      float nd = is_done ? 0.0f : 1.0f, dd = is_done ? 1.0f : 0.0f;
      float ret_info = *v.ep_ret * nd + new_ret * dd;
      int len_info = is_done ? new_len : *v.ep_len;
      *v.run_ret = new_ret * nd;
      *v.run_len = is_done ? 0 : new_len;
      *v.ep_ret = ret_info;
      *v.ep_len = len_info;
      done[e] = (uint8_t)is_done;
      ep_return[e] = ret_info;
      ep_length[e] = len_info;
      for (int i = 0; i < s->A; ++i) reward[(size_t)e * s->A + i] = rew;
      if (is_done && auto_reset) {
        uint32_t nk[2], unused[2];
        split2(v.key, nk, unused);
        generate(s, &v, nk);
        build_grids(s, &v, gsh, gag);
      }
      observe(s, &v, gsh, gag, view + (size_t)e * s->A * s->FR, mask + (size_t)e * s->A);
    }
    free(gsh);
    free(gag);
    free(niq);
  }
}
