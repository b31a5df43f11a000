// Backward of the FeaSt aggregation (training step, train_dual.py:199-218).
//
// Forward (feast.cu):  P = X U^T;  l_ijh = P_jh - P_ih + c_h;  q = softmax_h(l);  Z_i[h,:] = 1/d_i * sum_{j in N(i)+{i}} q_ijh x_j
// Given dZ [N, 9*C]:
//   dq_ijh  = 1/d_i * <dZ_i[h,:], x_j>                      dl_ijh = q_ijh (dq_ijh - sum_h' q_ijh' dq_ijh')
//   dx_j   += 1/d_i * sum_h q_ijh dZ_i[h,:]                 dP_jh += dl_ijh ;  dP_ih -= dl_ijh ;  dc_h += dl_ijh
// One warp per target node i (same traversal as the forward), lanes own channels; the per-edge 9 dot products are
// warp-reduced, scatter targets (dx_j, dP_j) are updated with red.global.add.f32.  The dense parts of the layer's backward
// (dZ = g.W_flat, dW = g^T.Z, dX += dP.U, dU = dP^T.X) are plain library GEMMs on the host side.
#include "common.cuh"

namespace geobi {

constexpr int H = GEOBI_HEADS;

template <int CPL>
__global__ void __launch_bounds__(256) feast_bwd_edges_kernel(const float* __restrict__ x, int64_t ldx, int64_t N, int C,
                                                              const int* __restrict__ rowptr, const int* __restrict__ nbr,
                                                              const double* __restrict__ P, const float* __restrict__ cvec,
                                                              const float* __restrict__ dZ, int64_t lddz, float* __restrict__ dx,
                                                              int64_t lddx, float* __restrict__ dP, float* __restrict__ dc) {
  __shared__ __align__(16) float qs[8][32][12];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 8 + warp;
  if (i >= N) return;
  const int b = rowptr[i];
  const int total = rowptr[i + 1] - b + 1;
  const float rcnt = 1.0f / (float)total;
  double Pi[H];
  float ch[H];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    Pi[h] = P[i * H + h];
    ch[h] = cvec[h];
  }
  // dZ_i slice owned by this lane: channels lane + 32k
  float dz[H][CPL];
#pragma unroll
  for (int h = 0; h < H; ++h)
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const int c = lane + 32 * k;
      dz[h][k] = c < C ? dZ[i * lddz + h * C + c] * rcnt : 0.f;
    }
  float dPi[H], dcl[H];
#pragma unroll
  for (int h = 0; h < H; ++h) dPi[h] = dcl[h] = 0.f;

  for (int s0 = 0; s0 < total; s0 += 32) {
    const int s = s0 + lane;
    int j = (int)i;
    if (s < total) {
      if (s > 0) j = nbr[b + s - 1];
      float l[H];
      float m = -INFINITY;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        l[h] = p_diff(P[(int64_t)j * H + h], Pi[h]) + ch[h];
        m = fmaxf(m, l[h]);
      }
      float sum = 0.f;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        l[h] = expf(l[h] - m);
        sum += l[h];
      }
#pragma unroll
      for (int h = 0; h < H; ++h) qs[warp][lane][h] = l[h] / sum;
    }
    __syncwarp();
    const int cnt = min(32, total - s0);
    for (int t = 0; t < cnt; ++t) {
      const int jt = __shfl_sync(0xffffffffu, j, t);
      float q[H], dq[H];
#pragma unroll
      for (int h = 0; h < H; ++h) {
        q[h] = qs[warp][t][h];
        dq[h] = 0.f;
      }
      float gx[CPL];
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        const int c = lane + 32 * k;
        const float xv = c < C ? x[(int64_t)jt * ldx + c] : 0.f;
        float acc = 0.f;
#pragma unroll
        for (int h = 0; h < H; ++h) {
          dq[h] = fmaf(dz[h][k], xv, dq[h]);
          acc = fmaf(q[h], dz[h][k], acc);
        }
        gx[k] = acc;
      }
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        const int c = lane + 32 * k;
        if (c < C) atomicAdd(dx + (int64_t)jt * lddx + c, gx[k]);
      }
      float dot = 0.f;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        dq[h] = warp_sum(dq[h]);
        dot = fmaf(q[h], dq[h], dot);
      }
      // lane h (< 9) finalises head h of this edge
      float dl = 0.f;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float v = q[h] * (dq[h] - dot);
        if (lane == h) dl = v;
        dPi[h] -= v;
        dcl[h] += v;
      }
      if (lane < H) atomicAdd(dP + (int64_t)jt * H + lane, dl);
    }
    __syncwarp();
  }
  // every lane holds identical dPi / dcl: lane h writes head h
#pragma unroll
  for (int h = 0; h < H; ++h)
    if (lane == h) {
      atomicAdd(dP + i * H + h, dPi[h]);
      atomicAdd(dc + h, dcl[h]);
    }
}

// dx[member argmax of segment s, c] = g[s, c]; every node belongs to exactly one segment, dx is zero-initialised.
__global__ void __launch_bounds__(256) segment_max_bwd_kernel(const float* __restrict__ x, int64_t ldx, int C, const int* __restrict__ rowptr,
                                                              const int* __restrict__ idx, int64_t n_seg, const float* __restrict__ g,
                                                              int64_t ldg, float* __restrict__ dx, int64_t lddx) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t s = t / C;
  if (s >= n_seg) return;
  const int c = (int)(t - s * C);
  const int b = rowptr[s], e = rowptr[s + 1];
  if (e <= b) return;
  int best = idx[b];
  float bv = x[(int64_t)best * ldx + c];
  for (int k = b + 1; k < e; ++k) {
    const int m = idx[k];
    const float v = x[(int64_t)m * ldx + c];
    if (v > bv) { bv = v; best = m; }   // ties: first (lowest id) member
  }
  dx[(int64_t)best * lddx + c] = g[s * ldg + c];
}

int feast_project_and_aggregate(const float* x, int64_t ldx, int64_t N, int c_in, const int32_t* rowptr, const int32_t* nbr,
                                const int32_t* row_map, int64_t n_src, const float* U, const float* c, double* P, void* Z, int64_t ldz,
                                int out_mode, cudaStream_t st);  // feast.cu
}  // namespace geobi

using namespace geobi;

// Z (fp32 [N, 9*C_in]) and P (fp64 [N, 9]) of the forward, for the backward pass.
extern "C" int geobi_feast_aggregate(const float* x, int64_t ldx, int64_t N, int c_in, const int32_t* rowptr, const int32_t* nbr,
                                     const float* U, const float* c, double* P, float* Z, void* stream) {
  GEOBI_REQUIRE(x && rowptr && U && c && P && Z && N >= 0 && c_in >= 1 && c_in <= 128, "feast_aggregate: bad arguments");
  if (N == 0) return GEOBI_OK;
  return feast_project_and_aggregate(x, ldx, N, c_in, rowptr, nbr, nullptr, N, U, c, P, Z, (int64_t)H * c_in, 0, static_cast<cudaStream_t>(stream));
}

extern "C" int geobi_feast_bwd_edges(const float* x, int64_t ldx, int64_t N, int c_in, const int32_t* rowptr, const int32_t* nbr,
                                     const double* P, const float* c, const float* dZ, float* dx, int64_t lddx, float* dP, float* dc,
                                     void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GEOBI_REQUIRE(x && rowptr && P && c && dZ && dx && dP && dc && N >= 0 && c_in >= 1 && c_in <= 128, "feast_bwd_edges: bad arguments");
  if (N == 0) return GEOBI_OK;
  const unsigned blocks = (unsigned)cdiv(N, 8);
  const int64_t lddz = (int64_t)H * c_in;
  if (c_in <= 32) feast_bwd_edges_kernel<1><<<blocks, 256, 0, st>>>(x, ldx, N, c_in, rowptr, nbr, P, c, dZ, lddz, dx, lddx, dP, dc);
  else if (c_in <= 64) feast_bwd_edges_kernel<2><<<blocks, 256, 0, st>>>(x, ldx, N, c_in, rowptr, nbr, P, c, dZ, lddz, dx, lddx, dP, dc);
  else feast_bwd_edges_kernel<4><<<blocks, 256, 0, st>>>(x, ldx, N, c_in, rowptr, nbr, P, c, dZ, lddz, dx, lddx, dP, dc);
  GEOBI_LAUNCH_OK("feast_bwd_edges");
  return GEOBI_OK;
}

extern "C" int geobi_segment_max_bwd(const float* x, int64_t ldx, int channels, const int32_t* rowptr, const int32_t* idx, int64_t n_seg,
                                     const float* g, int64_t ldg, float* dx, int64_t lddx, void* stream) {
  GEOBI_REQUIRE(x && rowptr && idx && g && dx && channels > 0 && n_seg >= 0, "segment_max_bwd: bad arguments");
  if (n_seg == 0) return GEOBI_OK;
  segment_max_bwd_kernel<<<(unsigned)cdiv(n_seg * channels, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ldx, channels, rowptr, idx,
                                                                                                               n_seg, g, ldg, dx, lddx);
  GEOBI_LAUNCH_OK("segment_max_bwd");
  return GEOBI_OK;
}
