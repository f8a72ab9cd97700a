// K3 / K3b: radius-r bilinear lookup into the correlation pyramid, all levels in one launch.
//
// Replaces MutliScalePairwise4DCorr.__call__ + bilinear_sampler (reference
// ezflow/similarity/correlation/pairwise.py:43-69, ezflow/utils/resampling.py:56-85) and the autograd
// adjoint of that op w.r.t. the pyramid.
//
// Every query (b,h,w) owns a private correlation map per level, so the lookup is a gather of one
// (2r+2)x(2r+2) patch per (query, level): the (2r+1)^2 window taps sit at integer offsets from the
// centre and therefore share one fractional part.  A block serves 32 consecutive queries:
//   phase 1  warps fetch whole patches from the warp-strip-major pyramid with 16-byte vector loads (a
//            patch row is 20 B in fp16: <= 3 chunks; 30 lanes fetch a full patch in one instruction),
//            dequantise and drop them into shared memory (fp32, odd per-query stride => conflict-free
//            in phase 2);
//   phase 2  lane = query, warps stride over the window taps; every store instruction writes 128
//            contiguous bytes of the (B, L*(2r+1)^2, H, W) output.
// HBM-bound: algorithmic bytes per query = L*[(2r+2)^2*es + (2r+1)^2*4] + 8  (SURVEY.md §8d).
#include "ezb_common.cuh"

#include <algorithm>
#include <climits>

#include <cstdlib>
#include <type_traits>

namespace ezb {
namespace k3 {

constexpr int QB = 32;        // queries per block
constexpr int NTHREADS = 256;
constexpr int MAX_RADIUS = 8;

struct LookupParams {
  const char* pyr;      // subtile-sequence pyramid (ezb_common.cuh)
  int H, W;             // level l is (H >> l) x (W >> l)
  int sw[PYR_MAX_LEVELS], sub_base[PYR_MAX_LEVELS];
  int n_tiles, n_groups;
  long long group_bytes;
  const float* deq;     // [B]
  const float* coords;  // (B,2,N)
  float* out;           // (B, L*K*K, N)
  const float* dout;    // coords-gradient mode: (B, L*K*K, N)
  float* dcoords;       // coords-gradient mode: (B,2,N)
  int B, N, L, r;
};

__device__ __forceinline__ int float_floor_to_int_clamped(float v, int lo, int hi) {
  // NaN -> lo; +-inf and huge values saturate
  float f = floorf(v);
  f = fminf(fmaxf(f, (float)lo), (float)hi);
  return (f == f) ? (int)f : lo;
}

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// 16-byte asynchronous global -> shared copy, L2-only (.cg); src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async_16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ int div3(int v) { return (v * 43691) >> 17; }  // exact for 0 <= v < 98304

// Gather geometry.  A window row is fetched as `ng` aligned 16-byte chunks (VW = 8 fp16 / 4 fp32 pixels of
// one subtile row); in shared memory a row keeps the ng*VW fetched pixels verbatim (storage dtype) and
// the (2r+2) taps start `off` = xs mod VW pixels in.
template <int ES>
struct Gather {
  static constexpr int VW = 16 / ES, VW_LOG = (ES == 2) ? 3 : 2;
  __host__ __device__ static constexpr int ng(int D) { return (VW - 1 + D + VW - 1) / VW; }
  __host__ __device__ static constexpr int row_el(int D) { return ng(D) * VW; }
  __device__ static __forceinline__ int chunk_floor(int xs) { return ((xs + 64) >> VW_LOG) - (64 >> VW_LOG); }  // xs >= -64
};

// DCOORDS = false: the lookup.  DCOORDS = true: its adjoint w.r.t. coords (autograd of grid_sample's grid argument,
// pairwise.py:50-67) — same gather, but phase 2 contracts the window's horizontal / vertical tap differences with dout
// instead of writing the taps; the per-warp partial sums of a query meet in shared memory.
template <typename T, bool DCOORDS>
__global__ void __launch_bounds__(NTHREADS, 3) corr_lookup_fwd_kernel(const LookupParams P, const int qstride_words) {
  extern __shared__ __align__(16) uint32_t sm_words[];  // [QB][qstride_words] patches, then [L*QB] int2 item parameters (+ [2][QB] sums)
  constexpr int ES = (int)sizeof(T);
  const int r = P.r, D = 2 * r + 2, K = 2 * r + 1, KK = K * K;
  const int b = blockIdx.y, grp = blockIdx.x, q0 = grp * QB;  // QB == PYR_QG: a block serves one query group
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* cx_ptr = P.coords + (long long)b * 2 * P.N;
  const float* cy_ptr = cx_ptr + P.N;
  const char* pair_base = P.pyr + (long long)b * P.n_tiles * P.n_groups * P.group_bytes + (long long)grp * P.group_bytes;
  const unsigned int tile_stride = (unsigned int)(P.n_groups * P.group_bytes);  // < 2^32, checked by the host
  int2* items = reinterpret_cast<int2*>(sm_words + QB * qstride_words);
  float* gsum = reinterpret_cast<float*>(items + P.L * QB);  // DCOORDS only
  if (DCOORDS && threadIdx.x < 2 * QB) gsum[threadIdx.x] = 0.f;

  // ---------------- phase 0: window origins ----------------
  // Every warp serves 4 queries of the block; lane k < 16 owns item (query k/4, level k%4) and keeps its
  // window origin in a register (shuffled to the other lanes below) and in shared memory (phase 2).
  using G = Gather<ES>;
  // (first chunk index << 20) | (chunks the window's 2r+2 columns actually span - 1) << 16 | (first row & 0xffff); chunk index and
  // row are small signed numbers.  A window spans ng chunks only when it starts in the last pixels of a chunk (fp16, r = 4: one
  // start in eight); the chunks beyond its last column are not fetched — each would be a DRAM line of its own
  int org_packed;
  {
    const int qi = (lane >> 2) & 3, l = lane & 3;
    const int q = min(q0 + warp * 4 + qi, P.N - 1);
    const float inv = 1.0f / (float)(1 << l);
    const int x0 = float_floor_to_int_clamped(__ldg(cx_ptr + q) * inv, -(1 << 24), 1 << 24);
    const int y0 = float_floor_to_int_clamped(__ldg(cy_ptr + q) * inv, -(1 << 24), 1 << 24);
    // clamped so that a window entirely outside the level stays entirely outside
    const int2 org = make_int2(max(-D, min(P.W >> l, x0 - r)), max(-D, min(P.H >> l, y0 - r)));
    if (lane < 16 && l < P.L) items[l * QB + warp * 4 + qi] = org;
    const int cf = G::chunk_floor(org.x);
    org_packed = (cf << 20) | ((G::chunk_floor(org.x + D - 1) - cf) << 16) | (org.y & 0xffff);  // span - 1 <= 15: any window that fits shared memory
  }

  // ---------------- phase 1: gather windows ----------------
  // task = (window row, chunk); a round = 32 tasks of one (level, query).  In fp16, 30 lanes fetch a whole
  // (2r+2)^2 window with one 16-byte asynchronous copy each (cp.async: global -> shared without a register
  // round trip), so a warp has the windows of its 4 queries x all levels (16 x 480 bytes) in flight at once.
  // Everything outside a level's valid area is stored as zero by the GEMM (zero operand rows), and chunks
  // outside the level are zero-filled here (src-size 0), so no per-element masking is needed.
  const int ng = G::ng(D), row_el = G::row_el(D);
  const int tasks = D * ng, rounds = (tasks + 31) >> 5;
  const uint32_t sm_base = (uint32_t)__cvta_generic_to_shared(sm_words);
  for (int rnd = 0; rnd < rounds; ++rnd) {
    const int t = rnd * 32 + lane;
    const int row = (ng == 3) ? div3(t) : (ng == 4) ? t >> 2 : t / ng;
    const int gi = t - row * ng;
    const bool task_ok = t < tasks;
    const uint32_t dst_task = sm_base + (row * row_el + gi * G::VW) * ES + warp * 4 * (qstride_words * 4);
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) {
      const int ql = warp * 4 + qi;
#pragma unroll
      for (int l = 0; l < PYR_MAX_LEVELS; ++l) {
        const int pk = __shfl_sync(0xffffffffu, org_packed, qi * 4 + l);
        // chunks beyond the window's last column are neither fetched nor zero-filled: phase 2 never reads them
        if (l < P.L && task_ok && gi <= ((pk >> 16) & 15)) {
          const int y = (int)(short)(pk & 0xffff) + row;
          const int xg = ((pk >> 20) + gi) * G::VW;  // first pixel of the chunk
          const char* src = pair_base;
          uint32_t nbytes = 0;
          if ((unsigned)y < (unsigned)(P.H >> l) && (unsigned)xg < (unsigned)(P.W >> l)) {
            const unsigned sidx = (unsigned)(P.sub_base[l] + (y >> 3) * P.sw[l] + (xg >> 3));
            if constexpr (ES == 2) {
              // an fp16 chunk is one whole row of one 8x8 subtile: section sidx % 4 of the group block, 16 bytes at
              // (y % 8) * 16 of the query's 128-byte line (ezb_common.cuh, pyr_record_offset with column % 8 == 0)
              src = pair_base + (unsigned long long)(sidx >> 2) * tile_stride + (((sidx & 3u) << 12) + ((unsigned)ql << 7) + ((unsigned)(y & 7) << 4));
            } else {
              const int col = (int)(sidx % PYR_SUBS_PER_TILE) * PYR_SUB + (y % PYR_SY) * PYR_SX + xg % PYR_SX;
              src = pair_base + (unsigned long long)(sidx / PYR_SUBS_PER_TILE) * tile_stride + (unsigned)pyr_record_offset(ES, ql, col);
            }
            nbytes = 16;
          }
          cp_async_16(dst_task + qi * (qstride_words * 4) + l * D * row_el * ES, src, nbytes);
        }
      }
    }
  }
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---------------- phase 2: bilinear taps, lane = query ----------------
  // work item = (level, i): the 2r+1 outputs with x-offset i, walked down the window so each new output
  // needs two fresh taps (the upper pair is the previous lower pair).  Stores are 128 contiguous bytes.
  // Every warp takes a contiguous run of items, so the per-level setup is done once per warp (with 4 levels
  // and 8 warps: two warps per level).
  const int q = q0 + lane;
  if constexpr (DCOORDS) {
    // d out / d x = deq * [(1-fy)*(v01-v00) + fy*(v11-v10)] / 2^l,  d out / d y = (bot - top) / 2^l with the forward's
    // horizontally interpolated rows `top`, `bot`; taps outside a level are the zeros the gather stored.
    float gx = 0.f, gy = 0.f;
    if (q < P.N) {
      const float deq = __ldg(P.deq + b);
      const float x = __ldg(cx_ptr + q), y = __ldg(cy_ptr + q);
      const float* dq = P.dout + (long long)b * P.L * KK * P.N + q;
      const T* sm_q = reinterpret_cast<const T*>(sm_words + lane * qstride_words);
      const int n_items = P.L * K, n_warps = NTHREADS / 32;
      for (int item = n_items * warp / n_warps; item < n_items * (warp + 1) / n_warps; ++item) {
        const int l = item / K, i = item - l * K;
        const float inv = 1.0f / (float)(1 << l);
        const float cx = x * inv, cy = y * inv;
        const float fx = cx - floorf(cx), fy = cy - floorf(cy);
        const int xs = items[l * QB + lane].x;
        const T* p = sm_q + l * D * row_el + (xs - G::chunk_floor(xs) * G::VW) + i;
        const float* d = dq + (long long)(l * KK + i * K) * P.N;
        float x0 = (float)p[0], x1 = (float)p[1];
        float top = x0 + fx * (x1 - x0), dtop = x1 - x0, sx = 0.f, sy = 0.f;
        for (int j = 0; j < K; ++j) {
          p += row_el;
          x0 = (float)p[0];
          x1 = (float)p[1];
          const float bot = x0 + fx * (x1 - x0), dbot = x1 - x0;
          const float g = __ldg(d);
          sx += g * (dtop + fy * (dbot - dtop));
          sy += g * (bot - top);
          top = bot;
          dtop = dbot;
          d += P.N;
        }
        gx += sx * (deq * inv);
        gy += sy * (deq * inv);
      }
    }
    atomicAdd(gsum + lane, gx);  // zeroed before the barrier that ends phase 1
    atomicAdd(gsum + QB + lane, gy);
    __syncthreads();
    if (threadIdx.x < 2 * QB) {
      const int c = threadIdx.x >> 5, qq = q0 + (threadIdx.x & 31);
      if (qq < P.N) P.dcoords[((long long)b * 2 + c) * P.N + qq] = gsum[threadIdx.x];
    }
    return;
  }
  if (q >= P.N) return;
  const float deq = __ldg(P.deq + b);
  const float x = __ldg(cx_ptr + q), y = __ldg(cy_ptr + q);
  float* out = P.out + (long long)b * P.L * KK * P.N + q;
  const T* sm_q = reinterpret_cast<const T*>(sm_words + lane * qstride_words);
  const int n_items = P.L * K, n_warps = NTHREADS / 32;
  int item = n_items * warp / n_warps;
  const int item_end = n_items * (warp + 1) / n_warps;
  while (item < item_end) {
    const int l = item / K, i0 = item - l * K, i1 = min(K, i0 + item_end - item);
    const float inv = 1.0f / (float)(1 << l);
    const float cx = x * inv, cy = y * inv;
    const float fx = cx - floorf(cx), fy = cy - floorf(cy);
    // a clamped patch origin means the window is entirely outside: the patch is all zeros, weights moot
    const float wl0 = deq * (1.f - fx), wl1 = deq * fx;
    const int xs = items[l * QB + lane].x;
    const int off = xs - G::chunk_floor(xs) * G::VW;  // taps start `off` pixels into the fetched row
    const T* pl = sm_q + l * D * row_el + off;
    float* o = out + (long long)(l * KK + i0 * K) * P.N;  // i offsets x (slow output index!), j offsets y   pairwise.py:53-61
    // NC adjacent window columns at a time: a window row's NC + 1 pixels give NC horizontally interpolated values (the
    // pixel between two columns is loaded once), two consecutive rows give the taps
    auto columns = [&](auto nc_tag, const T* p, float* oc) {
      constexpr int NC = decltype(nc_tag)::value;
      float top[NC], x[NC + 1];
#pragma unroll
      for (int a = 0; a <= NC; ++a) x[a] = (float)p[a];
#pragma unroll
      for (int a = 0; a < NC; ++a) top[a] = wl0 * x[a] + wl1 * x[a + 1];
      for (int j = 0; j < K; ++j) {
        p += row_el;
#pragma unroll
        for (int a = 0; a <= NC; ++a) x[a] = (float)p[a];
#pragma unroll
        for (int a = 0; a < NC; ++a) {
          const float bot = wl0 * x[a] + wl1 * x[a + 1];
          oc[(long long)a * K * P.N] = top[a] + fy * (bot - top[a]);
          top[a] = bot;
        }
        oc += P.N;
      }
    };
    int i = i0;
    for (; i + 3 <= i1; i += 3, o += 3ll * K * P.N) columns(std::integral_constant<int, 3>{}, pl + i, o);
    if (i + 2 <= i1) {
      columns(std::integral_constant<int, 2>{}, pl + i, o);
      i += 2;
      o += 2ll * K * P.N;
    }
    if (i < i1) columns(std::integral_constant<int, 1>{}, pl + i, o);
    item += i1 - i0;
  }
}

// ------------------------------------------------------------------------------------ K3b
struct LookupBwdParams {
  float* dlvl[EZB_MAX_LEVELS];
  long long qstride[EZB_MAX_LEVELS];
  int h[EZB_MAX_LEVELS], w[EZB_MAX_LEVELS], pitch[EZB_MAX_LEVELS];
  const float* dout;    // (B, L*K*K, N)
  const float* coords;  // (B,2,N)
  int B, N, L, r;
  int rmw_rows;         // plain vector read-modify-writes where a level's rows are 16-byte aligned (else atomics)
};

// dpyr[l][b,q,y,x] += sum over window taps of weight * dout.  A query's gradient window is private to that query (no
// inter-query conflicts).  The block stages dout for its 32 queries in shared memory (coalesced loads, lane = query);
// then a thread owns one window ROW of one (query, level): it slides along x keeping the vertically-combined value of
// the previous window column, so every gradient element costs two shared-memory reads, and adds the row to HBM as
// 8-byte-aligned pairs (red.global.add.v2.f32) with scalar ends.
__global__ void __launch_bounds__(NTHREADS) corr_lookup_bwd_kernel(const LookupBwdParams P) {
  extern __shared__ float sm[];
  const int r = P.r, D = 2 * r + 2, K = 2 * r + 1, KK = K * K;
  const int dstride = (P.L * KK) | 1;  // dout values per query (odd stride)
  float* s_d = sm;                     // [QB][dstride]
  const int b = blockIdx.y, q0 = blockIdx.x * QB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* cx_ptr = P.coords + (long long)b * 2 * P.N;
  const float* cy_ptr = cx_ptr + P.N;

  // load dout: lane = query (coalesced), warps stride over channels
  {
    const int q = q0 + lane;
    const float* src = P.dout + (long long)b * P.L * KK * P.N + q;
    const int nch = P.L * KK, nw = NTHREADS / 32;
    for (int c0 = warp; c0 < nch; c0 += 8 * nw) {  // 8 independent loads in flight per thread
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int c = c0 + u * nw;
        v[u] = (q < P.N && c < nch) ? __ldg(src + (long long)c * P.N) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int c = c0 + u * nw;
        if (c < nch) s_d[lane * dstride + c] = v[u];
      }
    }
  }
  __syncthreads();

  // thread -> (level, window row Y, query): consecutive threads take consecutive queries (conflict-free s_d reads)
  for (int idx = threadIdx.x; idx < QB * P.L * D; idx += NTHREADS) {
    const int ql = idx % QB, ly = idx / QB, l = ly / D, Y = ly - l * D;
    const int q = q0 + ql;
    if (q >= P.N) continue;
    const float inv = 1.0f / (float)(1 << l);
    const float cx = __ldg(cx_ptr + q) * inv, cy = __ldg(cy_ptr + q) * inv;
    const float fx = cx - floorf(cx), fy = cy - floorf(cy);
    const int hl = P.h[l], wl = P.w[l];
    const int xs = max(-D, min(wl, float_floor_to_int_clamped(cx, -(1 << 24), 1 << 24) - r));
    const int ys = max(-D, min(hl, float_floor_to_int_clamped(cy, -(1 << 24), 1 << 24) - r));
    const int yy = ys + Y;
    if (yy < 0 || yy >= hl) continue;
    // out(i,j) touches window (row j+a, col i+c) with weight wy[a]*wx[c]  =>  row Y collects j = Y (a=0) and j = Y-1 (a=1)
    const float wy0 = (Y < K) ? 1.f - fy : 0.f, wy1 = (Y >= 1) ? fy : 0.f;
    const float* d0 = s_d + ql * dstride + l * KK + min(Y, K - 1);  // + i*K: dout(i, j=Y)
    const float* d1 = s_d + ql * dstride + l * KK + max(Y - 1, 0);  // + i*K: dout(i, j=Y-1)
    float* row = P.dlvl[l] + ((long long)b * P.N + q) * P.qstride[l] + (long long)yy * P.pitch[l] + xs;
    float prev = 0.f;  // vertically combined dout of window column X-1
    auto next = [&](int X) {  // gradient of window element (Y, X); X must advance by one per call
      const float cur = (X < K) ? wy0 * d0[X * K] + wy1 * d1[X * K] : 0.f;
      const float g = (1.f - fx) * cur + fx * prev;
      prev = cur;
      return g;
    };
    if (P.rmw_rows && ((P.pitch[l] | (int)P.qstride[l]) & 3) == 0) {
      // The rows of this level start 16-byte aligned (pitch and per-query stride are multiples of 4 floats), and a map row of a
      // query is touched by exactly one thread per launch: plain 16-byte read-modify-writes of the aligned slots that
      // cover the window row (elements outside the window add 0) replace the atomics — the L2 atomic units were the limit.
      float* rowbase = row - xs;                 // element (yy, 0)
      const int xa = xs & ~3;                    // first aligned slot (floor, also for negative xs)
      float4 v[4];
      bool live[4];
#pragma unroll
      for (int sl = 0; sl < 4; ++sl) {
        const int x0 = xa + 4 * sl;
        float e[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int Xw = x0 + u - xs;            // window column of this element
          e[u] = (Xw >= 0 && Xw < D) ? next(Xw) : 0.f;
        }
        live[sl] = x0 >= 0 && x0 < wl && x0 - xs < D && x0 + 3 - xs >= 0;  // inside the image (whole slots: wl % 4 == 0) and touching the window
        v[sl] = make_float4(e[0], e[1], e[2], e[3]);
      }
      float4 old[4];
#pragma unroll
      for (int sl = 0; sl < 4; ++sl)
        if (live[sl]) old[sl] = *reinterpret_cast<const float4*>(rowbase + xa + 4 * sl);
#pragma unroll
      for (int sl = 0; sl < 4; ++sl)
        if (live[sl])
          *reinterpret_cast<float4*>(rowbase + xa + 4 * sl) =
              make_float4(old[sl].x + v[sl].x, old[sl].y + v[sl].y, old[sl].z + v[sl].z, old[sl].w + v[sl].w);
      continue;
    }
    int X = 0;
    if ((reinterpret_cast<uintptr_t>(row) >> 2) & 1) {  // first address 8-byte misaligned: scalar head
      const float g = next(0);
      if (xs >= 0 && xs < wl) atomicAdd(row, g);
      X = 1;
    }
    for (; X + 1 < D; X += 2) {
      const float g0 = next(X), g1 = next(X + 1);
      const int xx = xs + X;
      const bool v0 = xx >= 0 && xx < wl, v1 = xx + 1 >= 0 && xx + 1 < wl;
      if (v0 && v1) {
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(row + X), "f"(g0), "f"(g1) : "memory");
      } else if (v0) {
        atomicAdd(row + X, g0);
      } else if (v1) {
        atomicAdd(row + X + 1, g1);
      }
    }
    if (X < D) {
      const float g = next(X);
      if (xs + X >= 0 && xs + X < wl) atomicAdd(row + X, g);
    }
  }
}


// ------------------------------------------------------------------------------------ K3b, all lookups of a step at once
// RAFT runs 12 lookups per step through one pyramid (models/raft.py:114-129) and, once it converges, their windows of a query
// largely overlap: 12 launches of corr_lookup_bwd_kernel read-modify-write nearly the same DRAM lines 12 times (~1.3 lines per
// window row and launch: what bounds that kernel).  Here a block owns 8 queries for ALL lookups: per (query, level) a
// MT_H x MT_W fp32 tile in shared memory is placed where the lookups' windows cluster; every lookup's window
// gradients are added into the tile (same sliding-row arithmetic as above, dout staged per lookup), and the tile is added to
// HBM once, skipping untouched 16-byte slots.  The tile is centred on the window of the first lookup passed (the caller puts
// the most converged iterate there); a lookup whose window does not fit inside it (early iterations, scattered coordinates)
// takes the per-lookup path (atomics) inside the same kernel, so any coordinates are handled.
constexpr int MQB = 8, MT_H = 16, MT_W = 20, M_MAXLK = 16, M_MAXL = 4;
constexpr int M_NT = 320, M_NBUF = 3;  // one pass over 8 queries x 4 levels x 10 window rows; dout staged three lookups deep
constexpr int MT_STRIDE = MT_H * MT_W + 1;  // odd: the 8 queries of a quarter-warp update 8 different banks
struct LookupBwdMultiParams {
  float* dlvl[M_MAXL];
  long long qstride[M_MAXL];
  int h[M_MAXL], w[M_MAXL], pitch[M_MAXL];
  const float* dout[M_MAXLK];    // (B, L*K*K, N) each
  const float* coords[M_MAXLK];  // (B,2,N) each
  int n, B, N, L, r;
};

// RT: compile-time radius (RAFT: 4, small model 3) so that the window loops unroll; 0 = runtime radius
template <int RT>
__global__ void __launch_bounds__(M_NT, 3) corr_lookup_bwd_multi_kernel(const LookupBwdMultiParams P) {
  extern __shared__ __align__(16) float sm[];
  const int r = RT > 0 ? RT : P.r, D = 2 * r + 2, K = 2 * r + 1, KK = K * K, nch = P.L * KK;
  const int dstride = nch | 1;  // dout values per query (odd stride)
  const int dbuf = (MQB * dstride + 3) & ~3;                         // floats per dout buffer
  float* s_dall = sm;                                                // [M_NBUF][MQB][dstride]
  float* tiles = sm + M_NBUF * dbuf;                                 // [MQB * L][MT_STRIDE]
  int* s_org = reinterpret_cast<int*>(tiles + MQB * P.L * MT_STRIDE);  // [MQB * L][3]: ox, oy, any
  const int b = blockIdx.y, q0 = blockIdx.x * MQB, tid = threadIdx.x;

  auto window_origin = [&](int k, int q, int l, int& xs, int& ys, float& fx, float& fy) {
    const float inv = 1.0f / (float)(1 << l);
    const float* c = P.coords[k] + (long long)b * 2 * P.N;
    const float cx = __ldg(c + q) * inv, cy = __ldg(c + P.N + q) * inv;
    fx = cx - floorf(cx);
    fy = cy - floorf(cy);
    xs = max(-D, min(P.w[l], float_floor_to_int_clamped(cx, -(1 << 24), 1 << 24) - r));
    ys = max(-D, min(P.h[l], float_floor_to_int_clamped(cy, -(1 << 24), 1 << 24) - r));
    return xs + D > 0 && xs < P.w[l] && ys + D > 0 && ys < P.h[l];  // the window overlaps the level
  };

  // ---- phase 0: centre the tile of every (query, level) on the window of lookup 0 (the caller passes the most converged
  //      iterate first: the other late iterations cluster around it, the early ones take the per-lookup path); zero the tiles
  if (tid < MQB * P.L) {
    const int ql = tid % MQB, l = tid / MQB, q = q0 + ql;
    int xs = 0, ys = 0;
    float fx, fy;
    const bool any = q < P.N && P.n > 0 && window_origin(0, q, l, xs, ys, fx, fy);
    s_org[3 * tid] = any ? ((xs - (MT_W - D) / 2) & ~3) : INT_MIN / 2;  // 16-byte aligned columns (floor, also for negatives)
    s_org[3 * tid + 1] = any ? ys - (MT_H - D) / 2 : INT_MIN / 2;      // no tile: nothing fits
    s_org[3 * tid + 2] = any ? 1 : 0;
  }
  for (int i = tid; i < MQB * P.L * MT_STRIDE; i += M_NT) tiles[i] = 0.f;

  // dout of lookup k -> buffer k % M_NBUF: consecutive threads take consecutive queries (32-byte segments); asynchronous
  // 4-byte copies, ~8 per thread, issued two lookups ahead of their use
  auto stage = [&](int k) {
    float* s_d = s_dall + (k % M_NBUF) * dbuf;
    const float* src = P.dout[k] + (long long)b * nch * P.N + q0;
    const int ql = tid % MQB, c0 = tid / MQB, cstep = M_NT / MQB;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_d + ql * dstride);
    if (q0 + ql < P.N) {
      for (int c = c0; c < nch; c += cstep)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4u * c), "l"(src + (long long)c * P.N + ql) : "memory");
    } else {
      for (int c = c0; c < nch; c += cstep) s_d[ql * dstride + c] = 0.f;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage(0);
  if (P.n > 1) stage(1);
  for (int k = 0; k < P.n; ++k) {
    if (k + 1 < P.n) {
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();  // dout k has landed for every thread; everyone is done with lookup k - 1 (its buffer is refilled next)
    if (k + 2 < P.n) stage(k + 2);
    const float* s_d = s_dall + (k % M_NBUF) * dbuf;
    // ---- thread -> (level, window row Y, query)
    for (int idx = tid; idx < MQB * P.L * D; idx += M_NT) {
      const int ql = idx % MQB, ly = idx / MQB, l = ly / D, Y = ly - l * D;
      const int q = q0 + ql;
      if (q >= P.N) continue;
      int xs, ys;
      float fx, fy;
      if (!window_origin(k, q, l, xs, ys, fx, fy)) continue;
      const int hl = P.h[l], wl = P.w[l];
      const int yy = ys + Y;
      // out(i,j) touches window (row j+a, col i+c) with weight wy[a]*wx[c]  =>  row Y collects j = Y (a=0) and j = Y-1 (a=1)
      const float wy0 = (Y < K) ? 1.f - fy : 0.f, wy1 = (Y >= 1) ? fy : 0.f;
      const float* d0 = s_d + ql * dstride + l * KK + min(Y, K - 1);  // + i*K: dout(i, j=Y)
      const float* d1 = s_d + ql * dstride + l * KK + max(Y - 1, 0);  // + i*K: dout(i, j=Y-1)
      float prev = 0.f;  // vertically combined dout of window column X-1
      auto next = [&](int X) {  // gradient of window element (Y, X); X must advance by one per call
        const float cur = (X < K) ? wy0 * d0[X * K] + wy1 * d1[X * K] : 0.f;
        const float g = (1.f - fx) * cur + fx * prev;
        prev = cur;
        return g;
      };
      const int* org = s_org + 3 * (l * MQB + ql);
      if (xs >= org[0] && xs + D <= org[0] + MT_W && ys >= org[1] && ys + D <= org[1] + MT_H) {  // this lookup's window is inside the tile
        float* t = tiles + (l * MQB + ql) * MT_STRIDE + (yy - org[1]) * MT_W + (xs - org[0]);
        if constexpr (RT > 0) {
#pragma unroll
          for (int X = 0; X < 2 * RT + 2; ++X) t[X] += next(X);  // this (query, level, row) belongs to this thread: no conflicts
        } else {
          for (int X = 0; X < D; ++X) t[X] += next(X);
        }
      } else {
        if (yy < 0 || yy >= hl) continue;
        float* row = P.dlvl[l] + ((long long)b * P.N + q) * P.qstride[l] + (long long)yy * P.pitch[l] + xs;
        for (int X = 0; X < D; ++X) {
          const float g = next(X);
          if (xs + X >= 0 && xs + X < wl) atomicAdd(row + X, g);
        }
      }
    }
  }
  __syncthreads();
  // ---- flush: every touched 16-byte slot of every compact tile is added to HBM once
  constexpr int SLOTS = MT_W / 4;
  for (int idx = tid; idx < MQB * P.L * MT_H * SLOTS; idx += M_NT) {
    const int sl = idx % SLOTS, row = (idx / SLOTS) % MT_H, item = idx / (SLOTS * MT_H);
    const int ql = item % MQB, l = item / MQB, q = q0 + ql;
    const int* org = s_org + 3 * item;
    if (q >= P.N || !org[2]) continue;
    const int y = org[1] + row, x = org[0] + 4 * sl;
    if (y < 0 || y >= P.h[l] || x + 3 < 0 || x >= P.w[l]) continue;
    const float* tv = tiles + item * MT_STRIDE + row * MT_W + 4 * sl;
    const float4 v = make_float4(tv[0], tv[1], tv[2], tv[3]);
    if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) continue;
    float* dst = P.dlvl[l] + ((long long)b * P.N + q) * P.qstride[l] + (long long)y * P.pitch[l] + x;
    if (x >= 0 && x + 3 < P.w[l] && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      float4 o = *reinterpret_cast<const float4*>(dst);  // a map row of a query is touched by exactly one thread per launch
      o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
      *reinterpret_cast<float4*>(dst) = o;
    } else {
      const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (x + u >= 0 && x + u < P.w[l] && e[u] != 0.f) atomicAdd(dst + u, e[u]);
    }
  }
}

}  // namespace k3
}  // namespace ezb

using namespace ezb;

static int lookup_launch(const void* pyramid, const float* deq_scale, const float* coords, int B, int H, int W, int L, int radius,
                         int dtype, float* out, const float* dout, float* dcoords, void* stream) {
  const bool dc = dcoords != nullptr;
  EZB_REQUIRE(B > 0 && H > 0 && W > 0 && L >= 1, "bad shape B=%d H=%d W=%d L=%d", B, H, W, L);
  EZB_REQUIRE(radius >= 0 && radius <= k3::MAX_RADIUS, "corr_radius %d outside [0,%d]", radius, k3::MAX_RADIUS);
  EZB_REQUIRE(dtype == EZB_F32 || dtype == EZB_F16, "bad dtype %d", dtype);
  if (L > PYR_MAX_LEVELS) return fail(EZB_ERR_UNSUPPORTED, "num_levels %d > %d is not supported by the fused pyramid", L, PYR_MAX_LEVELS);
  PyrLayout lay;
  EZB_REQUIRE(pyr_layout(H, W, L, dtype, &lay), "feature map %dx%d too small for %d pyramid levels", H, W, L);
  EZB_REQUIRE((reinterpret_cast<uintptr_t>(pyramid) & 127) == 0, "pyramid buffer must be 128-byte aligned");
  EZB_REQUIRE((long long)lay.n_groups * lay.group_bytes < (1ll << 32), "feature map too large (H*W=%d)", H * W);
  k3::LookupParams P;
  const int N = H * W;
  P.H = H;
  P.W = W;
  P.pyr = static_cast<const char*>(pyramid);
  for (int l = 0; l < PYR_MAX_LEVELS; ++l) {
    P.sw[l] = lay.sw[l];
    P.sub_base[l] = lay.sub_base[l];
  }
  P.n_tiles = lay.n_tiles;
  P.n_groups = lay.n_groups;
  P.group_bytes = lay.group_bytes;
  P.deq = deq_scale;
  P.coords = coords;
  P.out = out;
  P.dout = dout;
  P.dcoords = dcoords;
  P.B = B;
  P.N = N;
  P.L = L;
  P.r = radius;
  const int D = 2 * radius + 2, es = lay.es;
  // per-query window words over all levels (must match Gather<ES> in the kernel)
  const int row_el = es == 2 ? k3::Gather<2>::row_el(D) : k3::Gather<4>::row_el(D);
  // per-query stride: an ODD number of 16-byte chunks (cp.async needs 16-byte alignment; an odd chunk count
  // spreads the queries over all 8 chunk positions of a 128-byte bank row for the lane = query reads)
  int qchunks = L * D * row_el * es / 16;
  qchunks |= 1;
  const int qwords = qchunks * 4;
  const size_t smem = (size_t)k3::QB * qwords * 4 + (size_t)L * k3::QB * sizeof(int2) + (dc ? 2 * k3::QB * sizeof(float) : 0);
  if (smem > 200 * 1024) return fail(EZB_ERR_UNSUPPORTED, "lookup window too large (L=%d, r=%d)", L, radius);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static_assert(k3::QB == PYR_QG, "one lookup block per pyramid query group");
  dim3 grid(lay.n_groups, B);
  auto launch = [&](auto kernel) -> int {
    EZB_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<grid, k3::NTHREADS, smem, st>>>(P, qwords);
    EZB_LAUNCH_OK();
    return EZB_OK;
  };
  if (dtype == EZB_F16) return dc ? launch(k3::corr_lookup_fwd_kernel<__half, true>) : launch(k3::corr_lookup_fwd_kernel<__half, false>);
  return dc ? launch(k3::corr_lookup_fwd_kernel<float, true>) : launch(k3::corr_lookup_fwd_kernel<float, false>);
}

extern "C" int ezb_corr_lookup_fwd(const void* pyramid, const float* deq_scale, const float* coords, int B, int H, int W, int L,
                                   int radius, int dtype, float* out, void* stream) {
  EZB_REQUIRE(pyramid && deq_scale && coords && out, "null pointer");
  return lookup_launch(pyramid, deq_scale, coords, B, H, W, L, radius, dtype, out, nullptr, nullptr, stream);
}

extern "C" int ezb_corr_lookup_bwd_coords(const void* pyramid, const float* deq_scale, const float* coords, const float* dout, int B,
                                          int H, int W, int L, int radius, int dtype, float* dcoords, void* stream) {
  EZB_REQUIRE(pyramid && deq_scale && coords && dout && dcoords, "null pointer");
  return lookup_launch(pyramid, deq_scale, coords, B, H, W, L, radius, dtype, nullptr, dout, dcoords, stream);
}

extern "C" int ezb_corr_grad_layout(int H, int W, int L, int* h, int* w, int* pitch, int64_t* qstride) {
  EZB_REQUIRE(H > 0 && W > 0 && L >= 1 && L <= EZB_MAX_LEVELS, "bad pyramid shape H=%d W=%d L=%d", H, W, L);
  LevelGeom g[EZB_MAX_LEVELS];
  EZB_REQUIRE(pyramid_geom(H, W, L, EZB_F32, g), "feature map %dx%d too small for %d pyramid levels", H, W, L);
  for (int l = 0; l < L; ++l) {
    if (h) h[l] = g[l].h;
    if (w) w[l] = g[l].w;
    if (pitch) pitch[l] = g[l].pitch;
    if (qstride) qstride[l] = g[l].qstride;
  }
  return EZB_OK;
}

extern "C" int ezb_corr_lookup_bwd(const float* dout, const float* coords, int B, int H, int W, int L, int radius,
                                   float* const* dlvl_ptrs, void* stream) {
  EZB_REQUIRE(dout && coords && dlvl_ptrs, "null pointer");
  EZB_REQUIRE(B > 0 && H > 0 && W > 0 && L >= 1 && L <= EZB_MAX_LEVELS, "bad shape B=%d H=%d W=%d L=%d", B, H, W, L);
  EZB_REQUIRE(radius >= 0 && radius <= k3::MAX_RADIUS, "corr_radius %d outside [0,%d]", radius, k3::MAX_RADIUS);
  LevelGeom g[EZB_MAX_LEVELS];
  EZB_REQUIRE(pyramid_geom(H, W, L, EZB_F32, g), "feature map %dx%d too small for %d pyramid levels", H, W, L);
  k3::LookupBwdParams P;
  const int N = H * W;
  for (int l = 0; l < L; ++l) {
    EZB_REQUIRE(dlvl_ptrs[l], "level %d gradient pointer is null", l);
    P.dlvl[l] = dlvl_ptrs[l];
    P.h[l] = g[l].h;
    P.w[l] = g[l].w;
    P.pitch[l] = g[l].pitch;
    P.qstride[l] = g[l].qstride;
  }
  P.dout = dout;
  P.coords = coords;
  P.B = B;
  P.N = N;
  P.L = L;
  P.r = radius;
  P.rmw_rows = (2 * radius + 2 <= 13 && getenv("EZB_LOOKUP_BWD_ATOMICS") == nullptr) ? 1 : 0;  // a window row spans at most 4 aligned slots
  for (int l = 0; l < L; ++l)
    if (reinterpret_cast<uintptr_t>(dlvl_ptrs[l]) & 15) P.rmw_rows = 0;
  const int K = 2 * radius + 1;
  const size_t smem = (size_t)k3::QB * ((L * K * K) | 1) * sizeof(float);
  if (smem > 200 * 1024) return fail(EZB_ERR_UNSUPPORTED, "lookup window too large (L=%d, r=%d)", L, radius);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EZB_CUDA_OK(cudaFuncSetAttribute(k3::corr_lookup_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k3::corr_lookup_bwd_kernel<<<dim3(cdiv(N, k3::QB), B), k3::NTHREADS, smem, st>>>(P);
  EZB_LAUNCH_OK();
  return EZB_OK;
}

extern "C" int ezb_corr_lookup_bwd_multi(const float* const* dout_ptrs, const float* const* coords_ptrs, int n_lookups, int B, int H, int W,
                                         int L, int radius, float* const* dlvl_ptrs, void* stream) {
  EZB_REQUIRE(dout_ptrs && coords_ptrs && dlvl_ptrs && n_lookups >= 0, "bad arguments");
  EZB_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && L >= 1 && L <= EZB_MAX_LEVELS, "bad shape B=%d H=%d W=%d L=%d", B, H, W, L);
  EZB_REQUIRE(radius >= 0 && radius <= k3::MAX_RADIUS, "corr_radius %d outside [0,%d]", radius, k3::MAX_RADIUS);
  const int K = 2 * radius + 1, D = K + 1;
  // tiles hold windows of at most 10 x 10 (radius <= 4) with room for a few pixels of drift; otherwise one launch per lookup
  if (L > k3::M_MAXL || D > 10 || getenv("EZB_LOOKUP_BWD_SEPARATE") != nullptr) {
    for (int k = 0; k < n_lookups; ++k)
      if (int rc = ezb_corr_lookup_bwd(dout_ptrs[k], coords_ptrs[k], B, H, W, L, radius, dlvl_ptrs, stream)) return rc;
    return EZB_OK;
  }
  LevelGeom g[EZB_MAX_LEVELS];
  EZB_REQUIRE(pyramid_geom(H, W, L, EZB_F32, g), "feature map %dx%d too small for %d pyramid levels", H, W, L);
  k3::LookupBwdMultiParams P;
  const int N = H * W;
  for (int l = 0; l < k3::M_MAXL; ++l) {
    if (l < L) EZB_REQUIRE(dlvl_ptrs[l], "level %d gradient pointer is null", l);
    P.dlvl[l] = l < L ? dlvl_ptrs[l] : nullptr;
    P.h[l] = l < L ? g[l].h : 0;
    P.w[l] = l < L ? g[l].w : 0;
    P.pitch[l] = l < L ? g[l].pitch : 0;
    P.qstride[l] = l < L ? g[l].qstride : 0;
  }
  P.B = B; P.N = N; P.L = L; P.r = radius;
  const size_t smem = ((size_t)k3::M_NBUF * ((k3::MQB * ((L * K * K) | 1) + 3) & ~3) + (size_t)k3::MQB * L * k3::MT_STRIDE) * sizeof(float) +
                      (size_t)k3::MQB * L * 3 * sizeof(int);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto kern = radius == 4 ? k3::corr_lookup_bwd_multi_kernel<4> : radius == 3 ? k3::corr_lookup_bwd_multi_kernel<3> : k3::corr_lookup_bwd_multi_kernel<0>;
  EZB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int k0 = 0; k0 < n_lookups; k0 += k3::M_MAXLK) {
    P.n = std::min(k3::M_MAXLK, n_lookups - k0);
    for (int k = 0; k < P.n; ++k) {
      EZB_REQUIRE(dout_ptrs[k0 + k] && coords_ptrs[k0 + k], "lookup %d: null pointer", k0 + k);
      P.dout[k] = dout_ptrs[k0 + k];
      P.coords[k] = coords_ptrs[k0 + k];
    }
    kern<<<dim3(cdiv(N, k3::MQB), B), k3::M_NT, smem, st>>>(P);
    EZB_LAUNCH_OK();
  }
  return EZB_OK;
}
