// dmvae_metrics.cu - validation metrics over bulk-generated waypoint trajectories.
//
// The reference judges its generator by comparing distributions of generated and human trajectories:
//   * the speed between consecutive waypoints (Distribution.calculate_human_velocities, Distribution.py:248-296), the
//     Jensen-Shannon divergence of the two speed histograms over 50 common edges (plot_velocity_distribution,
//     :309-331);
//   * how many trajectories visit each cell of a scenario grid (Spatial_Distribution._count_trajectories_per_grid,
//     Spatial_Distribution.py:387-431) and the RMSE between the two count maps (:434-493).
// It does so with Python loops over a few hundred trajectories.  For the 10^6 trajectories per scenario that the
// generation kernel produces, the per-trajectory parts become three HBM-bound scans (120 B read per trajectory and
// pass): speeds (+ running min / max), histogram over caller-supplied edges, per-trajectory cell counts.  The few
// dozen numbers that come out (49 counts, a count map) are turned into the divergence / RMSE on the host.
//
// Arithmetic: float32 inputs stay float32 as in the reference's NumPy scalar arithmetic (subtract, square, add, sqrt,
// divide, each rounded once; no fused multiply-add).  NumPy squares a float32 scalar through powf, which is not
// correctly rounded: a speed can differ from the reference's by one unit in the last place in ~0.1 % of the steps
// (tests/test_metrics_gpu.py states 3e-7 relative).  Binning and cell lookup are exact: comparisons against the same
// float64 edges, with an arithmetic first guess corrected by those comparisons.
#include "dmvae_common.cuh"
#include "dmvae_launch.h"

namespace dmvae {

struct TrajView {
  const float* p;   // (n, T, 3)
  long long n;
  int T;
  int ct, cx, cy;   // column of time, x, y: [t, x, y] (this library's generation output) or [x, y, t] (the reference's tracker order)
};

__device__ __forceinline__ bool step_speed(const TrajView& v, long long j, int s, float* out) {
  const float* a = v.p + ((size_t)j * v.T + s) * 3;
  const float dt = __fsub_rn(a[3 + v.ct], a[v.ct]);
  if (!(dt > 1e-6f)) return false;                       // Distribution.py:267: repeated (or decreasing) time stamps
  const float dx = __fsub_rn(a[3 + v.cx], a[v.cx]), dy = __fsub_rn(a[3 + v.cy], a[v.cy]);
  *out = __fdiv_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))), dt);
  return true;
}
// What the reference's sequential loop had appended last when it reached the first entry of trajectory j + 1: the
// value of the final entry of trajectory j, found by walking back over entries whose time step is degenerate.
__device__ float resolve_tail(const TrajView& v, long long j) {
  for (; j >= 0; --j)
    for (int k = v.T - 1; k >= 0; --k) {
      float s;
      if (step_speed(v, j, min(k, v.T - 2), &s)) return s;
    }
  return 0.f;   // nothing valid before: Distribution.py:276
}

// one thread per trajectory: T entries (entry k = step min(k, T - 2): the last point repeats the last step, :282-294)
__global__ void __launch_bounds__(256) speeds_kernel(const TrajView v, float* __restrict__ vel, unsigned int* __restrict__ minmax) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float lo = __int_as_float(0x7f800000), hi = 0.f;
  if (j < v.n) {
    float carry = 0.f;
    bool have = false;
    for (int k = 0; k < v.T; ++k) {
      float s;
      if (step_speed(v, j, min(k, v.T - 2), &s)) {
        carry = s;
      } else if (!have) {
        carry = resolve_tail(v, j - 1);
      }
      have = true;
      vel[(size_t)j * v.T + k] = carry;
      lo = fminf(lo, carry);
      hi = fmaxf(hi, carry);
    }
  }
  // speeds are >= 0: their bit patterns order like unsigned integers
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(minmax, __float_as_uint(lo));
    atomicMax(minmax + 1, __float_as_uint(hi));
  }
}

// Short trajectories (seq_len <= 16, the reference's 10 and 12): a block stages 128 consecutive trajectories in shared
// memory with coalesced 128-bit loads (a thread's own trajectory is 12 * seq_len bytes apart from its neighbour's: read
// directly, every load instruction of a warp would touch 32 different lines), writes the speeds back the same way, and
// issues two atomics per BLOCK for the running min / max.  Persistent blocks walk the tiles.
constexpr int SP_ROWS = 128, SP_MAXT = 16;   // 128 rows: the staged tile, the output tile and the cell counters stay under 48 KB of static shared memory
__global__ void __launch_bounds__(SP_ROWS) speeds_tile_kernel(const TrajView v, float* __restrict__ vel, unsigned int* __restrict__ minmax) {
  __shared__ __align__(16) float tile[SP_ROWS * SP_MAXT * 3];
  __shared__ __align__(16) float outb[SP_ROWS * SP_MAXT];
  __shared__ float wlo[SP_ROWS / 32], whi[SP_ROWS / 32];
  const int T = v.T, row_f = T * 3;
  float lo = __int_as_float(0x7f800000), hi = 0.f;
  const long long n_tiles = (v.n + SP_ROWS - 1) / SP_ROWS;
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const long long j0 = t * SP_ROWS;
    const int rows = (int)min((long long)SP_ROWS, v.n - j0);
    const int nf = rows * row_f;
    const float* src = v.p + (size_t)j0 * row_f;
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
      for (int i = threadIdx.x * 4; i < nf; i += SP_ROWS * 4) {
        if (i + 4 <= nf) *reinterpret_cast<float4*>(tile + i) = __ldg(reinterpret_cast<const float4*>(src + i));
        else for (int q = i; q < nf; ++q) tile[q] = __ldg(src + q);
      }
    } else {
      for (int i = threadIdx.x; i < nf; i += SP_ROWS) tile[i] = __ldg(src + i);
    }
    __syncthreads();
    const int r = threadIdx.x;
    if (r < rows) {
      const float* a = tile + r * row_f;
      float carry = 0.f;
      for (int k = 0; k < T; ++k) {
        const int s = min(k, T - 2);
        const float dt = __fsub_rn(a[3 * (s + 1) + v.ct], a[3 * s + v.ct]);
        if (dt > 1e-6f) {
          const float dx = __fsub_rn(a[3 * (s + 1) + v.cx], a[3 * s + v.cx]), dy = __fsub_rn(a[3 * (s + 1) + v.cy], a[3 * s + v.cy]);
          carry = __fdiv_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))), dt);
        } else if (k == 0) {
          carry = resolve_tail(v, j0 + r - 1);   // rare: the value the reference's loop had appended last
        }
        outb[r * T + k] = carry;
        lo = fminf(lo, carry);
        hi = fmaxf(hi, carry);
      }
    }
    __syncthreads();
    float* dst = vel + (size_t)j0 * T;
    const int no = rows * T;
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
      for (int i = threadIdx.x * 4; i < no; i += SP_ROWS * 4) {
        if (i + 4 <= no) *reinterpret_cast<float4*>(dst + i) = *reinterpret_cast<const float4*>(outb + i);
        else for (int q = i; q < no; ++q) dst[q] = outb[q];
      }
    } else {
      for (int i = threadIdx.x; i < no; i += SP_ROWS) dst[i] = outb[i];
    }
    __syncthreads();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { wlo[threadIdx.x >> 5] = lo; whi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < SP_ROWS / 32; ++w) { lo = fminf(lo, wlo[w]); hi = fmaxf(hi, whi[w]); }
    atomicMin(minmax, __float_as_uint(lo));      // speeds are >= 0: their bit patterns order like unsigned integers
    atomicMax(minmax + 1, __float_as_uint(hi));
  }
}

constexpr int HIST_MAX_BINS = 256;
struct HistEdges {
  double e[HIST_MAX_BINS + 1];
  int nb;
};
// np.histogram(values, bins=edges): bin i holds edges[i] <= x < edges[i + 1], the last bin also x == edges[nb];
// values outside [edges[0], edges[nb]] (and NaN) are not counted.
__global__ void __launch_bounds__(256) histogram_kernel(const float* __restrict__ x, long long m, const __grid_constant__ HistEdges h,
                                                         unsigned long long* __restrict__ counts) {
  // one counter per (bin, lane): the 32 lanes of a warp never hit the same shared-memory word however skewed the
  // distribution is (a plain per-bin counter serialises a warp whose values fall into a few bins); warps still share
  // the copies through atomics.  Up to 64 bins here, wider histograms use one counter per bin.
  __shared__ unsigned int local[64 * 32];
  const bool wide = h.nb > 64;
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) local[i] = 0u;
  __syncthreads();
  const double first = h.e[0], last = h.e[h.nb];
  const double scale = (double)h.nb / (last - first);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
    const double v = (double)x[i];
    if (!(v >= first && v <= last)) continue;
    int b = (int)((v - first) * scale);        // first guess (exact for uniform edges up to rounding) ...
    b = b < 0 ? 0 : (b >= h.nb ? h.nb - 1 : b);
    while (b > 0 && v < h.e[b]) --b;           // ... corrected by the comparisons np.histogram's search makes
    while (b < h.nb - 1 && v >= h.e[b + 1]) ++b;
    atomicAdd(&local[wide ? b : b * 32 + lane], 1u);
  }
  __syncthreads();
  if (wide) {
    for (int i = threadIdx.x; i < h.nb; i += blockDim.x)
      if (local[i]) atomicAdd(&counts[i], (unsigned long long)local[i]);
  } else {
    for (int i = threadIdx.x; i < h.nb; i += blockDim.x) {
      unsigned int t = 0;
      for (int l = 0; l < 32; ++l) t += local[i * 32 + ((l + i) & 31)];
      if (t) atomicAdd(&counts[i], (unsigned long long)t);
    }
  }
}

struct GridSpec {
  double x0, xstep, y0, ystep;   // edges: x0 + i * xstep, i < nx (np.arange), likewise y
  int nx, ny;                    // numbers of EDGES; the map is (ny - 1) x (nx - 1)
};
// np.clip(np.digitize(v, edges) - 1, 0, n_edges - 2): index of the last edge <= v, clipped into the map
__device__ __forceinline__ int cell_index(double v, double e0, double step, int n_edges) {
  if (!(v == v)) return n_edges - 2;             // NaN: digitize puts it past the last edge
  double g = floor((v - e0) * (1.0 / step));   // a guess: the comparisons below make it exact
  int i = g < -1.0 ? -1 : (g > (double)n_edges ? n_edges : (int)g);
  while (i >= 0 && (i >= n_edges || v < e0 + (double)i * step)) --i;
  while (i + 1 < n_edges && v >= e0 + (double)(i + 1) * step) ++i;
  return i < 0 ? 0 : (i > n_edges - 2 ? n_edges - 2 : i);
}
constexpr int GRID_SMEM_CELLS = 12032;   // 47 KB of per-block counters
// one thread per trajectory; a cell counts a trajectory once however many of its points fall into it
__global__ void __launch_bounds__(256) cells_kernel(const TrajView v, const GridSpec g, unsigned long long* __restrict__ counts) {
  __shared__ unsigned int local[GRID_SMEM_CELLS];
  const int w = g.nx - 1, cells = w * (g.ny - 1);
  const bool use_local = cells <= GRID_SMEM_CELLS;
  if (use_local) {
    for (int i = threadIdx.x; i < cells; i += blockDim.x) local[i] = 0u;
    __syncthreads();
  }
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < v.n; j += (long long)gridDim.x * blockDim.x) {
    const float* tp = v.p + (size_t)j * v.T * 3;
    for (int k = 0; k < v.T; ++k) {
      const int c = cell_index((double)tp[3 * k + v.cy], g.y0, g.ystep, g.ny) * w + cell_index((double)tp[3 * k + v.cx], g.x0, g.xstep, g.nx);
      bool seen = false;                       // visited by an earlier point of this trajectory?
      for (int q = 0; q < k && !seen; ++q)
        seen = c == cell_index((double)tp[3 * q + v.cy], g.y0, g.ystep, g.ny) * w + cell_index((double)tp[3 * q + v.cx], g.x0, g.xstep, g.nx);
      if (seen) continue;
      if (use_local) atomicAdd(&local[c], 1u);
      else atomicAdd(&counts[c], 1ull);
    }
  }
  if (use_local) {
    __syncthreads();
    for (int i = threadIdx.x; i < cells; i += blockDim.x)
      if (local[i]) atomicAdd(&counts[i], (unsigned long long)local[i]);
  }
}

// Short trajectories: same staging as speeds_tile_kernel; the cell of every point is computed once and kept in
// registers (compile-time indices), the "already visited by an earlier point" test is integer compares only.
__global__ void __launch_bounds__(SP_ROWS) cells_tile_kernel(const TrajView v, const GridSpec g, unsigned long long* __restrict__ counts) {
  __shared__ __align__(16) float tile[SP_ROWS * SP_MAXT * 3];
  __shared__ unsigned int local[4096];
  const int w = g.nx - 1, cells = w * (g.ny - 1);
  const bool use_local = cells <= 4096;
  const int T = v.T, row_f = T * 3;
  if (use_local) for (int i = threadIdx.x; i < cells; i += SP_ROWS) local[i] = 0u;
  const long long n_tiles = (v.n + SP_ROWS - 1) / SP_ROWS;
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const long long j0 = t * SP_ROWS;
    const int rows = (int)min((long long)SP_ROWS, v.n - j0);
    const int nf = rows * row_f;
    const float* src = v.p + (size_t)j0 * row_f;
    __syncthreads();
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
      for (int i = threadIdx.x * 4; i < nf; i += SP_ROWS * 4) {
        if (i + 4 <= nf) *reinterpret_cast<float4*>(tile + i) = __ldg(reinterpret_cast<const float4*>(src + i));
        else for (int q = i; q < nf; ++q) tile[q] = __ldg(src + q);
      }
    } else {
      for (int i = threadIdx.x; i < nf; i += SP_ROWS) tile[i] = __ldg(src + i);
    }
    __syncthreads();
    const int r = threadIdx.x;
    if (r < rows) {
      const float* a = tile + r * row_f;
      int c[SP_MAXT];
#pragma unroll
      for (int k = 0; k < SP_MAXT; ++k)
        c[k] = k < T ? cell_index((double)a[3 * k + v.cy], g.y0, g.ystep, g.ny) * w + cell_index((double)a[3 * k + v.cx], g.x0, g.xstep, g.nx) : -1;
#pragma unroll
      for (int k = 0; k < SP_MAXT; ++k) {
        if (k >= T) break;
        bool seen = false;
#pragma unroll
        for (int q = 0; q < k; ++q) seen = seen || c[q] == c[k];
        if (seen) continue;
        if (use_local) atomicAdd(&local[c[k]], 1u);
        else atomicAdd(&counts[c[k]], 1ull);
      }
    }
  }
  if (use_local) {
    __syncthreads();
    for (int i = threadIdx.x; i < cells; i += SP_ROWS)
      if (local[i]) atomicAdd(&counts[i], (unsigned long long)local[i]);
  }
}

static TrajView make_view(const float* traj, long long n, int T, int layout) {
  TrajView v;
  v.p = traj; v.n = n; v.T = T;
  if (layout == 0) { v.ct = 0; v.cx = 1; v.cy = 2; }   // [t, x, y]
  else { v.cx = 0; v.cy = 1; v.ct = 2; }               // [x, y, t]
  return v;
}

cudaError_t launch_speeds(const float* traj, long long n, int T, int layout, float* vel, float* minmax, cudaStream_t stream) {
  const unsigned int init[2] = {0x7f800000u, 0u};      // +inf, 0
  cudaError_t e = cudaMemcpyAsync(minmax, init, sizeof(init), cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return e;
  const long long blocks = (n + 255) / 256;
  if (T <= SP_MAXT) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles = (n + SP_ROWS - 1) / SP_ROWS;
    const long long grid = tiles < (long long)sms * 6 ? tiles : (long long)sms * 6;
    speeds_tile_kernel<<<(unsigned int)grid, SP_ROWS, 0, stream>>>(make_view(traj, n, T, layout), vel, reinterpret_cast<unsigned int*>(minmax));
  } else {
    speeds_kernel<<<(unsigned int)blocks, 256, 0, stream>>>(make_view(traj, n, T, layout), vel, reinterpret_cast<unsigned int*>(minmax));
  }
  return cudaGetLastError();
}

cudaError_t launch_histogram(const float* values, long long m, const double* edges, int nb, unsigned long long* counts, int sm_count,
                             cudaStream_t stream) {
  HistEdges h;
  h.nb = nb;
  for (int i = 0; i <= nb; ++i) h.e[i] = edges[i];
  cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)nb * sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return e;
  long long blocks = (m + 255) / 256;
  if (blocks > (long long)sm_count * 8) blocks = (long long)sm_count * 8;
  histogram_kernel<<<(unsigned int)blocks, 256, 0, stream>>>(values, m, h, counts);
  return cudaGetLastError();
}

cudaError_t launch_cells(const float* traj, long long n, int T, int layout, double x0, double xstep, int nx, double y0, double ystep,
                         int ny, unsigned long long* counts, int sm_count, cudaStream_t stream) {
  GridSpec g{x0, xstep, y0, ystep, nx, ny};
  cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)(nx - 1) * (ny - 1) * sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return e;
  long long blocks = (n + 255) / 256;
  if (blocks > (long long)sm_count * 4) blocks = (long long)sm_count * 4;
  long long tiles = (n + SP_ROWS - 1) / SP_ROWS;
  if (tiles > (long long)sm_count * 5) tiles = (long long)sm_count * 5;
  if (T <= SP_MAXT) cells_tile_kernel<<<(unsigned int)tiles, SP_ROWS, 0, stream>>>(make_view(traj, n, T, layout), g, counts);
  else cells_kernel<<<(unsigned int)blocks, 256, 0, stream>>>(make_view(traj, n, T, layout), g, counts);
  return cudaGetLastError();
}

}  // namespace dmvae
