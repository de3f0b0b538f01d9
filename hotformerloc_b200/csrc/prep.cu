// Pre-octree input transforms on the device (opt-in):
// Normalize (datasets/augmentation.py:212-235, bounding-box form), the range masks of the evaluation loop
// (eval/pnv_evaluate.py:163-167) and CylindricalCoordinates (datasets/coordinate_utils.py:68-116) for a whole
// batch of raw float32 clouds, followed by a stable per-cloud compaction -- the input of hfl_octree_build.
// Every operation is done in the reference's order and precision with explicit IEEE rounding (no FMA contraction):
// fp32 subtract / divide / multiply, the fp64 np.interp rescale (slope * (x - xp0) + fp0) and the final fp32 rounding
// are bit-identical to the host path (measured: the normalised and height coordinates agree bit for bit).  Two
// functions are NOT reproducible bit for bit, because torch's CPU kernels for them are not correctly rounded and
// depend on the host build: sqrt (MKL VML on the AVX-512 build here: 1 ulp low for 0.7 % of inputs; __fsqrt_rn is
// correctly rounded) and atan2 (vectorised Sleef vs CUDA libm: <= 2 ulp apart, identical for ~73 % of inputs).  A
// last-bit difference of a coordinate moves a point into the neighbouring octree cell with probability ~1e-6.  That
// is why the host path (which runs the very kernels the reference runs) stays the default and this one is a switch.
#include "common.cuh"

namespace hfl {

struct PrepParams {
  const float* in;          // [n_in, 3] raw points, clouds back to back
  const int32_t* off_in;    // [B + 1]
  float* tmp;               // [n_in, 3] per-cloud compacted (cloud b starts at off_in[b])
  int32_t* cnt;             // [B] surviving points per cloud
  int B;
  int norm;                 // 0: none, 1: bounding-box normalisation
  int zero_mean;
  float scale_factor;       // > 0: coords / scale_factor, else coords * (two_range / (max extent + 1e-6))
  float two_range;          // 2 * norm_range
  int cyl;
  double phi_slope;         // (1 - (-1)) / (pi - (-pi)) as numpy computes it
};

__device__ __forceinline__ float block_reduce_minmax(float v, bool is_max, float* sm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float w = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, w) : fminf(v, w);
  }
  __syncthreads();
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  float r = sm[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = is_max ? fmaxf(r, sm[i]) : fminf(r, sm[i]);
  return r;
}

__global__ void __launch_bounds__(256) k_prep_cloud(const PrepParams p) {
  __shared__ float sm[8];
  __shared__ int wsum[8];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t s = p.off_in[b], e = p.off_in[b + 1];
  float cx = 0.f, cy = 0.f, cz = 0.f, factor = 1.f;
  if (p.norm) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = s + tid; i < e; i += 256)
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float v = p.in[3 * i + a];
        lo[a] = fminf(lo[a], v);
        hi[a] = fmaxf(hi[a], v);
      }
    float ext = -INFINITY;
    float c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float l = block_reduce_minmax(lo[a], false, sm), h = block_reduce_minmax(hi[a], true, sm);
      c[a] = __fmul_rn(__fadd_rn(l, h), 0.5f);
      ext = fmaxf(ext, __fsub_rn(h, l));
    }
    if (p.zero_mean) { cx = c[0]; cy = c[1]; cz = c[2]; }
    if (!(p.scale_factor > 0.f)) factor = __fdiv_rn(p.two_range, __fadd_rn(ext, 1.0e-6f));
  }
  const double PI = 3.141592653589793;
  int running = 0;
  for (int64_t base = s; base < e; base += 256) {
    const int64_t i = base + tid;
    bool keep = i < e;
    float x = 0.f, y = 0.f, z = 0.f;
    if (keep) {
      x = p.in[3 * i]; y = p.in[3 * i + 1]; z = p.in[3 * i + 2];
      if (p.norm) {
        if (p.zero_mean) { x = __fsub_rn(x, cx); y = __fsub_rn(y, cy); z = __fsub_rn(z, cz); }
        if (p.scale_factor > 0.f) {
          x = __fdiv_rn(x, p.scale_factor); y = __fdiv_rn(y, p.scale_factor); z = __fdiv_rn(z, p.scale_factor);
        } else {
          x = __fmul_rn(x, factor); y = __fmul_rn(y, factor); z = __fmul_rn(z, factor);
        }
      }
      keep = fabsf(x) <= 1.0f && fabsf(y) <= 1.0f && fabsf(z) <= 1.0f;
      if (keep && p.cyl) {
        const float rho = __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));
        keep = rho <= 1.0f;
        if (keep) {
          const float phi = atan2f(y, x);
          // np.interp(rho, [0, 1], [-1, 1]) and np.interp(phi, [-pi, pi], [-1, 1]) in fp64, then fp32
          double rs = rho >= 1.0f ? 1.0 : __dadd_rn(__dmul_rn(2.0, __dsub_rn((double)rho, 0.0)), -1.0);
          const double ph = (double)phi;
          double ps;
          if (ph >= PI) ps = 1.0;
          else if (ph <= -PI) ps = -1.0;
          else ps = __dadd_rn(__dmul_rn(p.phi_slope, __dsub_rn(ph, -PI)), -1.0);
          x = fminf(fmaxf(__double2float_rn(rs), -1.0f), 1.0f);
          y = fminf(fmaxf(__double2float_rn(ps), -1.0f), 1.0f);
          z = fminf(fmaxf(z, -1.0f), 1.0f);
        }
      }
    }
    // stable compaction inside the cloud: block-wide exclusive scan of the keep flags
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const int c = wsum[w];
      if (w < warp) before += c;
      total += c;
    }
    if (keep) {
      const int64_t o = s + running + before + __popc(bal & ((1u << lane) - 1u));
      p.tmp[3 * o] = x; p.tmp[3 * o + 1] = y; p.tmp[3 * o + 2] = z;
    }
    running += total;
    __syncthreads();
  }
  if (tid == 0) p.cnt[b] = running;
}

__global__ void k_prep_offsets(const int32_t* __restrict__ cnt, int B, int32_t* __restrict__ off_out,
                               int32_t* __restrict__ total_out) {
  // B <= 32767: one thread block, serial over chunks of 1024 with a running carry
  __shared__ int sm[1024];
  int carry = 0;
  for (int base = 0; base < B; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < B ? cnt[i] : 0;
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int t = threadIdx.x >= o ? sm[threadIdx.x - o] : 0;
      __syncthreads();
      sm[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < B) off_out[i] = carry + sm[threadIdx.x] - v;
    carry += sm[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) { off_out[B] = carry; if (total_out) *total_out = carry; }
}

__global__ void k_prep_compact(const float* __restrict__ tmp, const int32_t* __restrict__ off_in,
                               const int32_t* __restrict__ cnt, const int32_t* __restrict__ off_out, int B,
                               int64_t n_in, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_in; i += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = B;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if ((int64_t)__ldg(off_in + mid) <= i) lo = mid; else hi = mid;
    }
    const int64_t local = i - off_in[lo];
    if (local < cnt[lo]) {
      const int64_t o = (int64_t)off_out[lo] + local;
      out[3 * o] = tmp[3 * i]; out[3 * o + 1] = tmp[3 * i + 1]; out[3 * o + 2] = tmp[3 * i + 2];
    }
  }
}

}  // namespace hfl

using namespace hfl;

extern "C" {

int hfl_prepare_clouds(const float* in, const int32_t* off_in, int32_t B, int64_t n_in, int32_t norm,
                       int32_t zero_mean, float scale_factor, float norm_range, int32_t cyl, float* tmp,
                       int32_t* cnt, float* out, int32_t* off_out, int32_t* total_out, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  HFL_CHECK_ARG(in && off_in && tmp && cnt && out && off_out, "null argument");
  HFL_CHECK_ARG(B >= 1 && B < 32768 && n_in >= 0 && n_in < (1ll << 31), "bad batch / point count");
  PrepParams p{in, off_in, tmp, cnt, B, norm, zero_mean, scale_factor, 2.0f * norm_range, cyl,
               (1.0 - (-1.0)) / (3.141592653589793 - (-3.141592653589793))};
  HFL_LAUNCH((k_prep_cloud<<<B, 256, 0, st>>>(p)));
  HFL_LAUNCH((k_prep_offsets<<<1, 1024, 0, st>>>(cnt, B, off_out, total_out)));
  if (n_in > 0) HFL_LAUNCH((k_prep_compact<<<grid_for(n_in, 256), 256, 0, st>>>(tmp, off_in, cnt, off_out, B, n_in, out)));
  return HFL_OK;
}

}  // extern "C"
